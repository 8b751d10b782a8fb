// gftt.cu -- a5/a6: cv.goodFeaturesToTrack (VP.py:503,707; LKF.py:488,691).
//   a5  cornerMinEigenVal: Sobel3 (float32, cv2's exact op order) -> dx^2, dxdy, dy^2 -> block x block box SUM in
//       FP64 -> lambda_min in unfused float32                       (oracle lko_min_eigen)
//   a6  masked max -> threshold -> 3x3 non-max suppression -> compaction (warp ballots) -> sort by
//       (value desc, raster index desc) -> greedy min-distance on a cell grid, in score order
//                                                                    (oracle lko_gftt_select)
#include "common.cuh"
#include <math.h>

namespace {

// ---------------------------------------------------------------------------------------------
// a5: min-eigenvalue map.  Output tile 64 x 16; u8 tile with (r0+1, r1+1) halo; products at the
// REFLECTED coordinate (a product outside the image is the product AT the mirrored pixel, not the
// Sobel of the mirrored image: dxdy changes sign and the dy row pass is not mirror-symmetric in fp32).
// ---------------------------------------------------------------------------------------------
constexpr int ME_W = 64, ME_H = 16, ME_NT = 256;

__global__ void __launch_bounds__(ME_NT)
k_min_eigen(const uint8_t *__restrict__ src, long long rp, long long ip, int H, int W, float *__restrict__ dst,
            long long drp, long long dip, int bs, int tail_cols, float s, float s2, const int32_t *__restrict__ enable)
{
    extern __shared__ __align__(16) unsigned char smem[];
    if (enable && !enable[blockIdx.z]) return;
    const int r0 = bs / 2;                        // window = [-r0, bs-1-r0]
    const int pw = ME_W + bs - 1, ph = ME_H + bs - 1;     // product tile
    const int uw = pw + 2, uh = ph + 2;                   // u8 tile
    const int uws = (uw + 3) & ~3;
    double *hs = reinterpret_cast<double *>(smem);                           // [3][ph][ME_W]
    float *prod = reinterpret_cast<float *>(hs + 3 * ph * ME_W);            // [3][ph][pw]
    uint8_t *u8 = reinterpret_cast<uint8_t *>(prod + 3 * ph * pw);          // [uh][uws]
    const int x0 = blockIdx.x * ME_W, y0 = blockIdx.y * ME_H, b = blockIdx.z;
    const uint8_t *img = src + (long long)b * ip;
    const int ux0 = x0 - r0 - 1, uy0 = y0 - r0 - 1;
    for (int i = threadIdx.x; i < uh * uw; i += ME_NT) {
        int r = i / uw, c = i - r * uw;
        u8[r * uws + c] = img[(long long)reflect101(uy0 + r, H) * rp + reflect101(ux0 + c, W)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ph * pw; i += ME_NT) {
        int r = i / pw, c = i - r * pw;
        // true coordinate of this product, and its position in the u8 tile
        int qx = reflect101(x0 - r0 + c, W), qy = reflect101(y0 - r0 + r, H);
        int tx = qx - ux0, ty = qy - uy0;
        // mirrored coordinates of far-out tile cells (beyond what any output needs) may leave the tile
        if (tx < 1 || tx > uw - 2 || ty < 1 || ty > uh - 2) { prod[i] = 0.f; prod[ph * pw + i] = 0.f; prod[2 * ph * pw + i] = 0.f; continue; }
        const uint8_t *p0 = u8 + (ty - 1) * uws + tx, *p1 = p0 + uws, *p2 = p1 + uws;
        // Dx: row pass r = I[x+1]-I[x-1] (exact), column pass fma(s, r[y-1]+r[y+1], s2*r[y])
        float ra = (float)((int)p0[1] - (int)p0[-1]), rb = (float)((int)p1[1] - (int)p1[-1]), rc = (float)((int)p2[1] - (int)p2[-1]);
        float dx = __fmaf_rn(s, __fadd_rn(ra, rc), __fmul_rn(s2, rb));
        // Dy: row pass smoothing (vector body fused, scalar tail unfused), column pass row[y+1]-row[y-1]
        float top, bot;
        if (qx < W - tail_cols) {
            top = __fmaf_rn(s, (float)p0[1], __fmaf_rn(s2, (float)p0[0], __fmul_rn(s, (float)p0[-1])));
            bot = __fmaf_rn(s, (float)p2[1], __fmaf_rn(s2, (float)p2[0], __fmul_rn(s, (float)p2[-1])));
        } else {
            top = __fadd_rn(__fadd_rn(__fmul_rn(s, (float)p0[-1]), __fmul_rn(s2, (float)p0[0])), __fmul_rn(s, (float)p0[1]));
            bot = __fadd_rn(__fadd_rn(__fmul_rn(s, (float)p2[-1]), __fmul_rn(s2, (float)p2[0])), __fmul_rn(s, (float)p2[1]));
        }
        float dy = __fsub_rn(bot, top);
        prod[i] = __fmul_rn(dx, dx);
        prod[ph * pw + i] = __fmul_rn(dx, dy);
        prod[2 * ph * pw + i] = __fmul_rn(dy, dy);
    }
    __syncthreads();
    // horizontal box sums in FP64
    for (int i = threadIdx.x; i < ph * ME_W; i += ME_NT) {
        int r = i / ME_W, c = i - r * ME_W;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float *p = prod + k * ph * pw + r * pw + c;
            double acc = 0.0;
            for (int j = 0; j < bs; j++) acc = __dadd_rn(acc, (double)p[j]);
            hs[k * ph * ME_W + i] = acc;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ME_H * ME_W; i += ME_NT) {
        int r = i / ME_W, c = i - r * ME_W;
        int x = x0 + c, y = y0 + r;
        if (x >= W || y >= H) continue;
        double sa = 0.0, sb = 0.0, sc = 0.0;
        for (int j = 0; j < bs; j++) {
            sa = __dadd_rn(sa, hs[(r + j) * ME_W + c]);
            sb = __dadd_rn(sb, hs[ph * ME_W + (r + j) * ME_W + c]);
            sc = __dadd_rn(sc, hs[2 * ph * ME_W + (r + j) * ME_W + c]);
        }
        float a = __fmul_rn((float)sa, 0.5f), bb = (float)sb, c2 = __fmul_rn((float)sc, 0.5f);
        float t = __fsub_rn(a, c2);
        float v = __fsub_rn(__fadd_rn(a, c2), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(bb, bb))));
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(dst) + (long long)b * dip + (long long)y * drp + 4LL * x) = v;
    }
}

__device__ __forceinline__ unsigned f2key(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // order-preserving; 0 is below every float
}
__device__ __forceinline__ float key2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Candidate histogram (per list) over the order-preserving keys, counted from the list's maximum downwards: bin width 2^16
// key units = 1/128 octave.  The selection kernel uses it to sort only the strongest candidates it can possibly need.
constexpr int GF_BINS = 2048, GF_BIN_SHIFT = 16;
__device__ __forceinline__ int key_bin(unsigned kmax, unsigned key)
{
    const unsigned d = (kmax - key) >> GF_BIN_SHIFT;
    return d >= (unsigned)GF_BINS ? GF_BINS - 1 : (int)d;
}

constexpr int MAXM = 8;      // masks per image handled in one pass

struct MaskSet {
    const uint8_t *p; long long rp, ip; int per_image;   // per_image: 1 -> plane index b*n_masks+m, 0 -> m
};

// ---------------------------------------------------------------------------------------------
// a5, streaming form for odd block sizes 3/5/7 on 4-byte aligned rows with W % 4 == 0 (860, 1920, 3840 ...): same
// arithmetic as k_min_eigen, but every value is computed once and nothing is staged through CTA-wide barriers.
//   * A WARP owns 120 output columns (lanes 1..30, 4 columns each; lanes 0 and 31 carry the 3-column halo) and
//     slides down a strip of rows.  Per image row a lane loads one aligned 32-bit word, gets the two neighbour
//     bytes by shuffle, keeps the Sobel row terms of the last two rows in registers and forms the three float32
//     products of its 4 columns (cv2's op order, explicitly rounded intrinsics).
//   * Horizontal box sums: neighbour products by shuffle, FP64 sliding sum over the lane's 4 columns.
//   * Vertical box sums: FP64 running sum V += H(y+R) - H(y-R-1); the last 8 rows of H live in a per-lane ring in
//     shared memory (a lane only ever touches its own slots: no barrier).  Every term is a float32 product, so
//     the FP64 sums are exact and order-free (7x7 terms span < 2^53 except for ulp-residue products, see A.4).
//   * Borders: a product outside the image is the product AT the mirrored pixel (see k_min_eigen): in x the edge
//     lanes mirror their own columns, in y the ring is addressed with the mirrored row.
// ---------------------------------------------------------------------------------------------
constexpr int MES_WARPS = 2, MES_COLS = 120, MES_PF = 3;
template <int BS> struct MES {
    static constexpr int R = BS / 2, RING = 2 * R + 1;
    static constexpr size_t SMEM_PER_WARP = (size_t)RING * 3 * 128 * sizeof(float);
    static constexpr size_t SMEM = MES_WARPS * SMEM_PER_WARP;
};

__device__ __forceinline__ float u8f(unsigned word, unsigned sel)
{   // exact u8 -> f32: 0x4B0000vv is 8388608 + vv
    return __uint_as_float(__byte_perm(word, 0x4B000000u, sel)) - 8388608.0f;
}
__device__ __forceinline__ double shfl_up_d(double v)
{
    return __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(v), 1), __shfl_up_sync(0xffffffffu, __double2loint(v), 1));
}
__device__ __forceinline__ double shfl_down_d(double v)
{
    return __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(v), 1), __shfl_down_sync(0xffffffffu, __double2loint(v), 1));
}

// EDGE: the warp holds the first and/or last column group of the image (mirror selects needed); interior warps skip them
template <int BS, bool EDGE>
__device__ __forceinline__ void
mes_body(unsigned char *mes_smem, const uint8_t *__restrict__ src, long long rp, long long ip, int H, int W, float *__restrict__ dst,
         long long drp, long long dip, int tail_cols, float s, float s2, int rows_per_strip, int dst_al16, int wx0,
         MaskSet ms, int n_masks, unsigned *__restrict__ maxkey)
{
    constexpr int R = MES<BS>::R, RING = MES<BS>::RING;
    const int b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cx = wx0 - 4 + 4 * lane;                      // first of this lane's 4 columns
    const bool valid = cx >= 0 && cx < W;
    const bool first = EDGE && cx == 0, last = EDGE && cx == W - 4;
    const bool emit = valid && lane >= 1 && lane <= 30;
    // ring[slot][ch][lane]: the float32 products of the last 2R+1 rows, one float4 per lane (conflict-free 128-bit access)
    float4 *ring = reinterpret_cast<float4 *>(mes_smem + (size_t)warp * MES<BS>::SMEM_PER_WARP) + lane;
    const uint8_t *img = src + (long long)b * ip + (valid ? cx : 0);
    const int y0 = blockIdx.y * rows_per_strip, y1 = min(H, y0 + rows_per_strip);
    bool tl[4];
#pragma unroll
    for (int j = 0; j < 4; j++) tl[j] = cx + j >= W - tail_cols;
    const bool any_tail = tail_cols > 0;

    auto load_word = [&](int y) -> uint32_t {               // y: any row index, mirrored here
        return valid ? __ldg(reinterpret_cast<const uint32_t *>(img + (long long)reflect101(y, H) * rp)) : 0u;
    };
    // Sobel row terms of one image row for the lane's 4 columns: rd = I[x+1]-I[x-1] (exact), rs = [1 2 1] smoothing * s
    auto row_terms = [&](uint32_t w, float (&rd)[4], float (&rs)[4]) {
        const uint32_t wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
        float f[6];
        f[0] = first ? u8f(w, 0x7651) : u8f(wl, 0x7653);    // column -1 -> 1
        f[1] = u8f(w, 0x7650); f[2] = u8f(w, 0x7651); f[3] = u8f(w, 0x7652); f[4] = u8f(w, 0x7653);
        f[5] = last ? u8f(w, 0x7652) : u8f(wr, 0x7650);     // column W -> W-2
#pragma unroll
        for (int j = 0; j < 4; j++) {
            rd[j] = __fsub_rn(f[j + 2], f[j]);
            rs[j] = __fmaf_rn(s, f[j + 2], __fmaf_rn(s2, f[j + 1], __fmul_rn(s, f[j])));
            if (any_tail && tl[j])
                rs[j] = __fadd_rn(__fadd_rn(__fmul_rn(s, f[j]), __fmul_rn(s2, f[j + 1])), __fmul_rn(s, f[j + 2]));
        }
    };
    // horizontal FP64 box sums of the lane's 4 columns of one plane of vertical sums (neighbour columns by shuffle)
    auto hbox = [&](const double (&p)[4], double (&h)[4]) {
        double v[10];
#pragma unroll
        for (int j = 0; j < 4; j++) v[3 + j] = p[j];
#pragma unroll
        for (int k = 1; k <= R; k++) {
            const double l = shfl_up_d(v[3 + 4 - k]);          // column -k   (left lane's column 4-k)
            const double r = shfl_down_d(v[3 + k - 1]);        // column 3+k  (right lane's column k-1)
            v[3 - k] = first ? v[3 + k] : l;                   // column -k -> k
            v[6 + k] = last ? v[3 + 3 - k] : r;                // column W-1+k -> W-1-k
        }
        double acc = v[3 - R];
#pragma unroll
        for (int k = 4 - R; k <= 3 + R; k++) acc = __dadd_rn(acc, v[k]);
        h[0] = acc;
#pragma unroll
        for (int j = 1; j < 4; j++) {
            acc = __dsub_rn(__dadd_rn(acc, v[3 + R + j]), v[2 - R + j]);
            h[j] = acc;
        }
    };
    auto ring_ld = [&](int row, int ch, float (&p)[4]) {
        const float4 q = ring[((row % RING) * 3 + ch) * 32];
        p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w;
    };
    auto ring_st = [&](int row, int ch, const float (&p)[4]) { ring[((row % RING) * 3 + ch) * 32] = make_float4(p[0], p[1], p[2], p[3]); };
    // fused a6 step 1: masked maximum of the map per (image, mask) as order-preserving keys (values are >= 0)
    unsigned best[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; m++) best[m] = 0;
    auto emit_row = [&](int y, const double (&vxx)[4], const double (&vxy)[4], const double (&vyy)[4]) {
        if (!emit) return;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float a = __fmul_rn((float)vxx[j], 0.5f), bb = (float)vxy[j], c2 = __fmul_rn((float)vyy[j], 0.5f);
            const float t = __fsub_rn(a, c2);
            o[j] = __fsub_rn(__fadd_rn(a, c2), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(bb, bb))));
        }
        float *q = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(dst) + (long long)b * dip + (long long)y * drp) + cx;
        if (dst_al16) *reinterpret_cast<float4 *>(q) = make_float4(o[0], o[1], o[2], o[3]);
        else { q[0] = o[0]; q[1] = o[1]; q[2] = o[2]; q[3] = o[3]; }
        if (maxkey) {
            unsigned k[4];
#pragma unroll
            for (int j = 0; j < 4; j++) k[j] = f2key(o[j]);
            if (!ms.p) best[0] = max(best[0], max(max(k[0], k[1]), max(k[2], k[3])));
            else {
#pragma unroll
                for (int m = 0; m < MAXM; m++) {
                    if (m >= n_masks) break;
                    const uint8_t *mp = ms.p + (long long)(ms.per_image ? b * n_masks + m : m) * ms.ip + (long long)y * ms.rp + cx;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (mp[j]) best[m] = max(best[m], k[j]);
                }
            }
        }
    };

    // product rows ylo..yhi are computed in order; image rows ylo-1 .. yhi+1 (mirrored) feed them, prefetched MES_PF ahead
    const int ylo = max(0, y0 - R), yhi = min(H - 1, y1 - 1 + R);
    float rd_m[4], rs_m[4], rd_c[4], rs_c[4];
    row_terms(load_word(ylo - 1), rd_m, rs_m);
    row_terms(load_word(ylo), rd_c, rs_c);
    uint32_t pf[MES_PF];
#pragma unroll
    for (int k = 0; k < MES_PF; k++) pf[k] = load_word(min(ylo + 1 + k, yhi + 1));

    // Vertical FP64 running sums of the products per column (every term is a float32 product: the sums are exact and
    // order-free), then the horizontal sums of those.  V of an output row: the first time the direct (mirrored) window sum
    // from the ring, afterwards + entering row - leaving row.
    double vxx[4], vxy[4], vyy[4];
    bool have_v = false;
    int ynext = y0;
    auto advance = [&](int y, const float (&nxx)[4], const float (&nxy)[4], const float (&nyy)[4],
                       const float (&oxx)[4], const float (&oxy)[4], const float (&oyy)[4]) {
        if (!have_v) {
#pragma unroll
            for (int j = 0; j < 4; j++) { vxx[j] = 0.0; vxy[j] = 0.0; vyy[j] = 0.0; }
#pragma unroll
            for (int k = -R; k <= R; k++) {
                float qxx[4], qxy[4], qyy[4];
                const int row = reflect101(y + k, H);
                ring_ld(row, 0, qxx); ring_ld(row, 1, qxy); ring_ld(row, 2, qyy);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    vxx[j] = __dadd_rn(vxx[j], (double)qxx[j]); vxy[j] = __dadd_rn(vxy[j], (double)qxy[j]);
                    vyy[j] = __dadd_rn(vyy[j], (double)qyy[j]);
                }
            }
            have_v = true;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                vxx[j] = __dsub_rn(__dadd_rn(vxx[j], (double)nxx[j]), (double)oxx[j]);
                vxy[j] = __dsub_rn(__dadd_rn(vxy[j], (double)nxy[j]), (double)oxy[j]);
                vyy[j] = __dsub_rn(__dadd_rn(vyy[j], (double)nyy[j]), (double)oyy[j]);
            }
        }
        double hxx[4], hxy[4], hyy[4];
        hbox(vxx, hxx); hbox(vxy, hxy); hbox(vyy, hyy);
        emit_row(y, hxx, hxy, hyy);
    };

#pragma unroll 1
    for (int r = ylo; r <= yhi; r++) {
        const uint32_t w = pf[0];
#pragma unroll
        for (int k = 0; k + 1 < MES_PF; k++) pf[k] = pf[k + 1];
        pf[MES_PF - 1] = load_word(min(r + 1 + MES_PF, yhi + 1));
        float rd_p[4], rs_p[4];
        row_terms(w, rd_p, rs_p);
        float pxx[4], pxy[4], pyy[4], oxx[4], oxy[4], oyy[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float dx = __fmaf_rn(s, __fadd_rn(rd_m[j], rd_p[j]), __fmul_rn(s2, rd_c[j]));
            const float dy = __fsub_rn(rs_p[j], rs_m[j]);
            pxx[j] = __fmul_rn(dx, dx); pxy[j] = __fmul_rn(dx, dy); pyy[j] = __fmul_rn(dy, dy);
            rd_m[j] = rd_c[j]; rd_c[j] = rd_p[j]; rs_m[j] = rs_c[j]; rs_c[j] = rs_p[j];
        }
        const bool out = r - R == ynext;                    // warp-uniform
        if (out && have_v) {                                // leaving row: read before its slot is overwritten (RING = 2R+1)
            const int ro = reflect101(ynext - R - 1, H);
            ring_ld(ro, 0, oxx); ring_ld(ro, 1, oxy); ring_ld(ro, 2, oyy);
        }
        ring_st(r, 0, pxx); ring_st(r, 1, pxy); ring_st(r, 2, pyy);
        if (out) { advance(ynext, pxx, pxy, pyy, oxx, oxy, oyy); ynext++; }
    }
    // bottom of the image: the entering rows are mirrored rows that are still in the ring
    while (ynext < y1) {
        float nxx[4], nxy[4], nyy[4], oxx[4], oxy[4], oyy[4];
        if (have_v) {
            const int rn = reflect101(ynext + R, H), ro = reflect101(ynext - R - 1, H);
            ring_ld(rn, 0, nxx); ring_ld(rn, 1, nxy); ring_ld(rn, 2, nyy);
            ring_ld(ro, 0, oxx); ring_ld(ro, 1, oxy); ring_ld(ro, 2, oyy);
        }
        advance(ynext, nxx, nxy, nyy, oxx, oxy, oyy);
        ynext++;
    }
    if (maxkey) {
#pragma unroll
        for (int m = 0; m < MAXM; m++) {
            if (m >= n_masks) break;
            const unsigned v = __reduce_max_sync(0xffffffffu, best[m]);
            if (lane == 0 && v) atomicMax(&maxkey[b * n_masks + m], v);
        }
    }
}

template <int BS>
__global__ void __launch_bounds__(MES_WARPS * 32, 8)
k_min_eigen_stream(const uint8_t *__restrict__ src, long long rp, long long ip, int H, int W, float *__restrict__ dst,
                   long long drp, long long dip, int tail_cols, float s, float s2, int rows_per_strip, int dst_al16,
                   const int32_t *__restrict__ enable, MaskSet ms, int n_masks, unsigned *__restrict__ maxkey)
{
    extern __shared__ __align__(16) unsigned char mes_smem[];
    if (enable && !enable[blockIdx.z]) return;
    const int wx0 = (blockIdx.x * MES_WARPS + (threadIdx.x >> 5)) * MES_COLS;
    if (wx0 >= W) return;                                   // warp-uniform; the kernel has no CTA-wide barrier
    if (wx0 == 0 || wx0 + MES_COLS + 4 >= W)
        mes_body<BS, true>(mes_smem, src, rp, ip, H, W, dst, drp, dip, tail_cols, s, s2, rows_per_strip, dst_al16, wx0, ms, n_masks, maxkey);
    else
        mes_body<BS, false>(mes_smem, src, rp, ip, H, W, dst, drp, dip, tail_cols, s, s2, rows_per_strip, dst_al16, wx0, ms, n_masks, maxkey);
}

template <int BS>
int launch_min_eigen_stream(lkf_ctx *ctx, const lkf_image *src, const lkf_image *eig, int tail, float s, float s2,
                            const int32_t *enable, MaskSet ms, int n_masks, unsigned *maxkey, cudaStream_t st)
{
    LKF_FUNC_ATTRS({
        LKF_CUDA(cudaFuncSetAttribute(k_min_eigen_stream<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MES<BS>::SMEM));
        LKF_CUDA(cudaFuncSetAttribute(k_min_eigen_stream<BS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    });
    const int W = src->width, H = src->height, B = src->batch;
    const int warps_x = (W + MES_COLS - 1) / MES_COLS, ctas_x = (warps_x + MES_WARPS - 1) / MES_WARPS;
    // strip height: fill whole waves of resident CTAs (shared memory bounds residency), 8..64 rows per strip
    const int sms = ctx ? ctx->sm_count : 148;
    int per_sm = (int)((227 * 1024) / (MES<BS>::SMEM + 1024));
    per_sm = per_sm > 8 ? 8 : per_sm;                         // registers: <= 8 CTAs of 64 threads
    const long long cap = (long long)sms * per_sm, col_ctas = (long long)ctas_x * B;
    const long long min_ctas = col_ctas * ((H + 63) / 64);
    const long long waves = (min_ctas + cap - 1) / cap;
    long long strips = waves * cap / col_ctas;
    if (strips < 1) strips = 1;
    int rows = (int)((H + strips - 1) / strips);
    rows = rows < 8 ? 8 : (rows > 64 ? 64 : rows);
    dim3 grid(ctas_x, (H + rows - 1) / rows, B);
    k_min_eigen_stream<BS><<<grid, MES_WARPS * 32, MES<BS>::SMEM, st>>>((const uint8_t *)src->data, src->row_pitch, src->image_pitch, H, W,
                                                                        (float *)eig->data, eig->row_pitch, eig->image_pitch, tail, s, s2,
                                                                        rows, aligned_to(eig, 16) ? 1 : 0, enable, ms, n_masks, maxkey);
    return LKF_OK;
}

// ---------------------------------------------------------------------------------------------
// a6 helpers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_masked_max(const float *__restrict__ eig, long long erp, long long eip, int H, int W, MaskSet ms, int n_masks,
             unsigned *__restrict__ maxkey, const int32_t *__restrict__ enable)
{
    const int b = blockIdx.y;
    if (enable && !enable[b]) return;
    const unsigned char *e = reinterpret_cast<const unsigned char *>(eig) + (long long)b * eip;
    unsigned best[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; m++) best[m] = 0;
    const long long P = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        int y = (int)(i / W), x = (int)(i - (long long)y * W);
        unsigned k = f2key(*reinterpret_cast<const float *>(e + (long long)y * erp + 4LL * x));
#pragma unroll
        for (int m = 0; m < MAXM; m++) {
            if (m >= n_masks) break;
            bool on = true;
            if (ms.p) on = ms.p[(long long)(ms.per_image ? b * n_masks + m : m) * ms.ip + (long long)y * ms.rp + x] != 0;
            if (on && k > best[m]) best[m] = k;
        }
    }
#pragma unroll
    for (int m = 0; m < MAXM; m++) {
        if (m >= n_masks) break;
        unsigned v = __reduce_max_sync(0xffffffffu, best[m]);
        if ((threadIdx.x & 31) == 0 && v) atomicMax(&maxkey[b * n_masks + m], v);
    }
}

// threshold + 3x3 NMS + compaction: one candidate list per (image, mask)
__global__ void __launch_bounds__(256)
k_nms_compact(const float *__restrict__ eig, long long erp, long long eip, int H, int W, MaskSet ms, int n_masks,
              const unsigned *__restrict__ maxkey, double quality, unsigned long long *__restrict__ keys,
              int *__restrict__ counts, int cap, const int32_t *__restrict__ enable, int *__restrict__ hist)
{
    const int b = blockIdx.z;
    if (enable && !enable[b]) return;
    const int x = 1 + blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = 1 + blockIdx.y * 8 + (threadIdx.x >> 5);
    const unsigned char *e = reinterpret_cast<const unsigned char *>(eig) + (long long)b * eip;
    bool inside = x < W - 1 && y < H - 1;
    // per-mask thresholds; the 3x3 neighbourhood is only fetched where the pixel survives the weakest of them
    float thr_m[MAXM], thr_min = 3.4e38f;
#pragma unroll
    for (int m = 0; m < MAXM; m++) {
        if (m >= n_masks) break;
        const unsigned mk = maxkey[b * n_masks + m];
        thr_m[m] = (float)((mk ? (double)key2f(mk) : 0.0) * quality);
        thr_min = fminf(thr_min, thr_m[m]);
    }
    float v = 0.f;
    bool is_max = false;
    if (inside) {
        const float *r1 = reinterpret_cast<const float *>(e + (long long)y * erp) + x;
        v = r1[0];
        if (v > thr_min && v != 0.f) {
            const float *r0 = reinterpret_cast<const float *>(e + (long long)(y - 1) * erp) + x;
            const float *r2 = reinterpret_cast<const float *>(e + (long long)(y + 1) * erp) + x;
            float m = fmaxf(fmaxf(fmaxf(r0[-1], r0[0]), fmaxf(r0[1], r1[-1])), fmaxf(fmaxf(r1[1], r2[-1]), fmaxf(r2[0], r2[1])));
            is_max = v >= m;
        }
    }
    if (!__any_sync(0xffffffffu, is_max)) return;
#pragma unroll
    for (int m = 0; m < MAXM; m++) {
        if (m >= n_masks) break;
        const int list = b * n_masks + m;
        const float thr = thr_m[m];
        bool cand = inside && is_max && v > thr;          // v > thr >= ... ; thresholded neighbours are <= v
        if (cand && ms.p) cand = ms.p[(long long)(ms.per_image ? list : m) * ms.ip + (long long)y * ms.rp + x] != 0;
        // a thresholded map keeps v only if v > thr; "v != 0" then holds unless thr < 0 and v == 0
        if (cand && v == 0.f) cand = false;
        unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (bal) {
            int lane = threadIdx.x & 31, leader = __ffs(bal) - 1, base = 0;
            if (lane == leader) base = atomicAdd(&counts[list], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (cand) {
                int pos = base + __popc(bal & ((1u << lane) - 1));
                if (pos < cap)
                    keys[(long long)list * cap + pos] = ((unsigned long long)f2key(v) << 32) | (unsigned)(y * W + x);
                atomicAdd(&hist[list * GF_BINS + key_bin(maxkey[list], f2key(v))], 1);
            }
        }
    }
}

// Same selection, 4 pixels per thread (128-bit loads of the centre row; rows above / below only where a pixel survives the
// weakest threshold).  Needs 16-byte aligned eig rows.  Candidate order inside a list is irrelevant (sorted afterwards).
// The candidates of a CTA (128 x 32 pixels) are counted in shared memory and appended with ONE global atomicAdd per list
// (one per warp and row cost 47 % of the kernel's stall samples: ~30 000 same-address atomics per list).
constexpr int NMS4_ROWS = 4;        // rows per thread (their centre loads are issued together: memory-level parallelism)

__global__ void __launch_bounds__(256, 4)
k_nms_compact4(const float *__restrict__ eig, long long erp, long long eip, int H, int W, MaskSet ms, int n_masks,
               const unsigned *__restrict__ maxkey, double quality, unsigned long long *__restrict__ keys,
               int *__restrict__ counts, int cap, const int32_t *__restrict__ enable, int *__restrict__ hist)
{
    __shared__ int s_cnt[MAXM], s_base[MAXM];
    const int b = blockIdx.z;
    if (enable && !enable[b]) return;
    const int lane = threadIdx.x & 31;
    const int x0 = (blockIdx.x * 32 + lane) * 4;
    const int ybase = 1 + blockIdx.y * (8 * NMS4_ROWS) + (threadIdx.x >> 5);
    const unsigned char *e = reinterpret_cast<const unsigned char *>(eig) + (long long)b * eip;
    if (threadIdx.x < MAXM) s_cnt[threadIdx.x] = 0;
    float thr_m[MAXM], thr_min = 3.4e38f;
    unsigned mk[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; m++) {
        if (m >= n_masks) break;
        mk[m] = maxkey[b * n_masks + m];
        thr_m[m] = (float)((mk[m] ? (double)key2f(mk[m]) : 0.0) * quality);
        thr_min = fminf(thr_min, thr_m[m]);
    }
    const bool full = x0 + 3 < W;
    float4 cq[NMS4_ROWS];
#pragma unroll
    for (int r = 0; r < NMS4_ROWS; r++) {
        const int y = ybase + 8 * r;
        cq[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y < H - 1 && full) cq[r] = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(e + (long long)y * erp) + x0);
    }
    unsigned mx = 0;                     // bit 4 r + j: pixel (x0 + j, ybase + 8 r) is a local maximum above the weakest threshold
#pragma unroll
    for (int r = 0; r < NMS4_ROWS; r++) {
        const int y = ybase + 8 * r;
        if (y < H - 1 && x0 < W) {
            const float *r1 = reinterpret_cast<const float *>(e + (long long)y * erp) + x0;
            float c[6];                                     // columns x0-1 .. x0+4 of the centre row
            if (!full) {
                float t[4];
                for (int j = 0; j < 4; j++) t[j] = x0 + j < W ? r1[j] : 0.f;
                cq[r] = make_float4(t[0], t[1], t[2], t[3]);
            }
            c[1] = cq[r].x; c[2] = cq[r].y; c[3] = cq[r].z; c[4] = cq[r].w;
            bool any = false;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int x = x0 + j;
                any |= x >= 1 && x < W - 1 && c[1 + j] > thr_min && c[1 + j] != 0.f;
            }
            if (any) {
                const float *r0 = reinterpret_cast<const float *>(e + (long long)(y - 1) * erp) + x0;
                const float *r2 = reinterpret_cast<const float *>(e + (long long)(y + 1) * erp) + x0;
                float a[6], d[6];
                if (full) {
                    const float4 qa = *reinterpret_cast<const float4 *>(r0), qd = *reinterpret_cast<const float4 *>(r2);
                    a[1] = qa.x; a[2] = qa.y; a[3] = qa.z; a[4] = qa.w; d[1] = qd.x; d[2] = qd.y; d[3] = qd.z; d[4] = qd.w;
                } else
                    for (int j = 0; j < 4; j++) { a[1 + j] = x0 + j < W ? r0[j] : 0.f; d[1 + j] = x0 + j < W ? r2[j] : 0.f; }
                const bool hl = x0 > 0, hr = x0 + 4 < W;
                a[0] = hl ? r0[-1] : 0.f; c[0] = hl ? r1[-1] : 0.f; d[0] = hl ? r2[-1] : 0.f;
                a[5] = hr ? r0[4] : 0.f; c[5] = hr ? r1[4] : 0.f; d[5] = hr ? r2[4] : 0.f;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int x = x0 + j;
                    const float v = c[1 + j];
                    if (x >= 1 && x < W - 1 && v > thr_min && v != 0.f) {
                        const float m = fmaxf(fmaxf(fmaxf(a[j], a[j + 1]), fmaxf(a[j + 2], c[j])), fmaxf(fmaxf(c[j + 2], d[j]), fmaxf(d[j + 1], d[j + 2])));
                        if (v >= m) mx |= 1u << (4 * r + j);
                    }
                }
            }
        }
    }
    if (!__syncthreads_or(mx != 0)) return;               // (also orders the zeroing of s_cnt)
    // (few threads hold a maximum: the per-list work below loops over the set bits only)
    auto value = [&](int bit) -> float {
        const int r = bit >> 2, j = bit & 3;
        const float4 q = r == 0 ? cq[0] : (r == 1 ? cq[1] : (r == 2 ? cq[2] : cq[3]));
        return j == 0 ? q.x : (j == 1 ? q.y : (j == 2 ? q.z : q.w));
    };
    static_assert(NMS4_ROWS == 4, "value() selects among four rows");
    // per list: the pixels that pass its threshold and mask; slots inside the CTA's block of the list
    unsigned cb[MAXM];
    int ofs[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; m++) {
        if (m >= n_masks) break;
        cb[m] = 0; ofs[m] = 0;
        if (mx) {
            const int list = b * n_masks + m;
            for (unsigned rest = mx; rest; rest &= rest - 1) {
                const int bit = __ffs(rest) - 1;
                bool cand = value(bit) > thr_m[m];
                if (cand && ms.p) {
                    const int y = ybase + 8 * (bit >> 2), x = x0 + (bit & 3);
                    cand = ms.p[(long long)(ms.per_image ? list : m) * ms.ip + (long long)y * ms.rp + x] != 0;
                }
                if (cand) cb[m] |= 1u << bit;
            }
            if (cb[m]) ofs[m] = atomicAdd(&s_cnt[m], __popc(cb[m]));
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < n_masks && s_cnt[threadIdx.x] > 0)
        s_base[threadIdx.x] = atomicAdd(&counts[b * n_masks + threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MAXM; m++) {
        if (m >= n_masks) break;
        if (!cb[m]) continue;
        const int list = b * n_masks + m;
        int pos = s_base[m] + ofs[m];
        for (unsigned rest = cb[m]; rest; rest &= rest - 1) {
            const int bit = __ffs(rest) - 1;
            const int y = ybase + 8 * (bit >> 2), x = x0 + (bit & 3);
            const unsigned k = f2key(value(bit));
            if (pos < cap) keys[(long long)list * cap + pos] = ((unsigned long long)k << 32) | (unsigned)(y * W + x);
            atomicAdd(&hist[list * GF_BINS + key_bin(mk[m], k)], 1);
            pos++;
        }
    }
}

// sort (bitonic, one CTA per list) + min-distance selection, SURVEY A.5 (5)-(7).
//
// cv2's greedy pass ("visit candidates in score order; accept unless an already ACCEPTED corner lies
// within minDistance") is evaluated by a parallel fixed point that yields exactly the same set:
// candidate i is REJECTED as soon as a higher-ranked neighbour (distance < minDistance) is ACCEPTED, and
// ACCEPTED once all its higher-ranked neighbours are REJECTED; the strongest candidate has no
// higher-ranked neighbour, so by induction over the rank every candidate gets the greedy verdict.  All
// undecided candidates are examined concurrently each round (1024 threads); rounds = longest dependency
// chain (a handful on images).  Neighbours come from a uniform grid of linked lists (cell side >=
// minDistance, 3x3 search) held in shared memory; "stop at maxCorners" is a rank-ordered prefix count.
// (cv2 searches the 3x3 block of its round(minDistance) cells; a corner closer than minDistance is never
// two such cells away, so "any accepted corner within minDistance" is the same predicate.)
constexpr int SS_NT = 1024;
constexpr int SS_SMEM_KEYS = 8192;
constexpr int SS_MAX_CELLS = 16 * 1024;           // uint32 heads in shared memory (64 KB)
constexpr unsigned SS_NIL = 0xffffffffu;

__global__ void __launch_bounds__(SS_NT)
k_sort_select(unsigned long long *__restrict__ keys_all, const int *__restrict__ counts, int cap, int W, int H,
              int max_corners, double min_distance, int cs, int gw, int gh, unsigned *__restrict__ gnext_all,
              unsigned char *__restrict__ gstate_all, float *__restrict__ corners, int32_t *__restrict__ out_counts,
              int max_out, const int32_t *__restrict__ enable, int n_masks, const unsigned *__restrict__ maxkey,
              const int *__restrict__ hist_all, unsigned long long *__restrict__ gtop_all, const int *__restrict__ done)
{
    extern __shared__ __align__(16) unsigned char ss_smem[];
    __shared__ int s_scan[SS_NT / 32];
    __shared__ int s_base, s_cut, s_ntop, s_cnt;
    const int list = blockIdx.x, tid = threadIdx.x;
    if (enable && !enable[list / n_masks]) return;
    if (done && done[list]) return;                        // k_sort_select_fast has written this list
    unsigned long long *gkeys = keys_all + (long long)list * cap;
    const int n_all = counts[list];
    const bool overflow = n_all > cap;
    const int n_full = overflow ? cap : n_all;
    const int stop_at = (max_corners > 0 && max_corners <= max_out) ? max_corners : 0x7fffffff;

    // Strongest-first prefilter: the greedy pass never looks past the candidate that completes maxCorners, so only the
    // top `need` candidates (by histogram bin) are sorted first; the rare list that needs more is redone in full.
    if (tid == 0) { s_cut = -1; s_ntop = n_full; }
    __syncthreads();
    if (hist_all && !overflow && stop_at != 0x7fffffff) {
        const int need = stop_at + stop_at / 2 + 256;
        if (n_full > need + need / 4) {
            const int *hist = hist_all + list * GF_BINS;
            const int h0 = hist[2 * tid], h1 = hist[2 * tid + 1];
            int incl = h0 + h1;
            const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) s_scan[wid] = incl;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < wid; w++) woff += s_scan[w];
            incl += woff;
            const int excl = incl - h0 - h1;
            if (excl < need && need <= incl) {
                if (excl + h0 >= need) { s_cut = 2 * tid; s_ntop = excl + h0; }
                else { s_cut = 2 * tid + 1; s_ntop = incl; }
            }
            __syncthreads();
        }
    }
    const int cut = s_cut, n_top = s_ntop;
    const unsigned kmax = maxkey ? maxkey[list] : 0u;

    int n = n_full;
    unsigned *xy = nullptr;
    volatile unsigned char *state = nullptr;
    for (int attempt = 0; attempt < 2; attempt++) {
        const bool trunc = attempt == 0 && cut >= 0 && n_top < n_full;
        n = trunc ? n_top : n_full;
        int npad = 32;
        while (npad < n) npad <<= 1;
        const bool in_smem = npad <= SS_SMEM_KEYS;
        unsigned long long *keys = in_smem ? reinterpret_cast<unsigned long long *>(ss_smem)
                                           : (trunc ? gtop_all + (long long)list * cap : gkeys);
        unsigned *heads = reinterpret_cast<unsigned *>(ss_smem + (size_t)SS_SMEM_KEYS * 8);
        unsigned *nxt = in_smem ? heads + SS_MAX_CELLS : gnext_all + (size_t)list * cap;
        state = in_smem ? reinterpret_cast<unsigned char *>(heads + SS_MAX_CELLS + SS_SMEM_KEYS) : gstate_all + (size_t)list * cap;
        __syncthreads();
        if (trunc) {
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            for (int i = tid; i < n_full; i += SS_NT) {
                const unsigned long long k = gkeys[i];
                if (key_bin(kmax, (unsigned)(k >> 32)) <= cut) keys[atomicAdd(&s_cnt, 1)] = k;
            }
        } else if (in_smem) {
            for (int i = tid; i < n; i += SS_NT) keys[i] = gkeys[i];
        }
        for (int i = n + tid; i < npad; i += SS_NT) keys[i] = 0ull;
        __syncthreads();
        for (int k = 2; k <= npad; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (npad >> 1); t += SS_NT) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));      // lower index of the pair
                    int ixj = i | j;
                    unsigned long long a = keys[i], b = keys[ixj];
                    bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
                __syncthreads();
            }
        // low word of every key: raster index -> (y << 16 | x)
        xy = reinterpret_cast<unsigned *>(keys);                     // xy[2*i]
        for (int i = tid; i < n; i += SS_NT) {
            unsigned idx = xy[2 * i];
            unsigned y = idx / (unsigned)W, x = idx - y * (unsigned)W;
            xy[2 * i] = (y << 16) | x;
            state[i] = (min_distance >= 1) ? 0 : 1;
        }
        int accepted = n;
        if (min_distance >= 1) {
            for (int i = tid; i < gw * gh; i += SS_NT) heads[i] = SS_NIL;
            __syncthreads();
            for (int i = tid; i < n; i += SS_NT) {
                unsigned v = xy[2 * i];
                int cell = ((v >> 16) / cs) * gw + (v & 0xffff) / cs;
                nxt[i] = atomicExch(&heads[cell], (unsigned)i);
            }
            __syncthreads();
            const float md2 = (float)(min_distance * min_distance);     // integer coordinates: d2 is exact
            const bool md2_exact = (double)md2 == min_distance * min_distance;
            // Rank-ordered windows of SS_NT candidates: every higher-ranked window is final, so only dependencies inside
            // the window need fixed-point rounds (a few), and the scan stops as soon as maxCorners corners are accepted.
            accepted = 0;
            for (int base = 0; base < n && accepted < stop_at; base += SS_NT) {
                const int i = base + tid;
                for (;;) {
                    int pending = 0;
                    if (i < n && !state[i]) {
                        unsigned v = xy[2 * i];
                        int x = v & 0xffff, y = v >> 16;
                        int cx = x / cs, cy = y / cs;
                        int verdict = 1;
                        for (int yy = max(0, cy - 1); yy <= min(gh - 1, cy + 1) && verdict != 2; yy++)
                            for (int xx = max(0, cx - 1); xx <= min(gw - 1, cx + 1) && verdict != 2; xx++)
                                for (unsigned j = heads[yy * gw + xx]; j != SS_NIL; j = nxt[j]) {
                                    if ((int)j >= i) continue;
                                    unsigned u = xy[2 * j];
                                    float dx = (float)(x - (int)(u & 0xffff)), dy = (float)(y - (int)(u >> 16));
                                    float d2 = dx * dx + dy * dy;
                                    bool close = md2_exact ? (d2 < md2) : ((double)d2 < min_distance * min_distance);
                                    if (close) {
                                        unsigned sj = state[j];
                                        if (sj == 1) { verdict = 2; break; }
                                        if (sj == 0) verdict = 0;
                                    }
                                }
                        if (verdict) state[i] = (unsigned char)verdict; else pending = 1;
                    }
                    if (!__syncthreads_or(pending)) break;
                }
                accepted += __syncthreads_count(i < n && state[i] == 1);
            }
        }
        if (!trunc || accepted >= stop_at) break;            // (block-uniform) otherwise: the full list, once
    }
    __syncthreads();
    // rank-ordered compaction of the accepted candidates, cut at `limit`
    float *out = corners + (size_t)list * max_out * 2;
    const int limit = (max_corners > 0 && max_corners < max_out) ? max_corners : max_out;
    const bool may_truncate = max_corners <= 0 || max_corners > max_out;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SS_NT) {
        int i = base + tid;
        int acc = (i < n && state[i] == 1) ? 1 : 0;
        unsigned bal = __ballot_sync(0xffffffffu, acc);
        int lane = tid & 31, wid = tid >> 5;
        if (lane == 0) s_scan[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
        for (int w = 0; w < SS_NT / 32; w++) { int c = s_scan[w]; if (w < wid) woff += c; total += c; }
        int pos = s_base + woff + __popc(bal & ((1u << lane) - 1));
        if (acc && pos < limit) {
            unsigned v = xy[2 * i];
            out[2 * pos] = (float)(v & 0xffff); out[2 * pos + 1] = (float)(v >> 16);
        }
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
        if (s_base >= limit && !may_truncate) break;
    }
    if (tid == 0) {
        int total = s_base;
        int nc = total < limit ? total : limit;
        bool truncated = may_truncate && total > limit;
        // negative count = capacity problem (candidate scratch or output rows too small): host raises
        out_counts[list] = (overflow || truncated) ? -1 - nc : nc;
    }
}


// ---------------------------------------------------------------------------------------------
// Fast selection path (the common case: maxCorners given, at most FS_MAX_N candidates to rank).
//   k_top_filter (many CTAs per list): the strongest candidates -- histogram bins 0..cut, cut chosen so that at least
//     `need` = 1.5 maxCorners + 256 candidates are kept -- are copied to a dense array; a CTA appends its matches with one
//     global atomicAdd.  (One CTA per list streaming 60 000 keys was a third of the old kernel's time.)
//   k_sort_select_fast (one CTA per list): RANKS the kept candidates instead of sorting them with a bitonic network:
//     counting sort by histogram bin (the bin prefix is known from the histogram), then the exact rank inside a bin by
//     counting the larger keys of that bin (bins hold a handful of candidates; a guard sends skewed lists to the general
//     kernel).  Keys are unique (the raster index is part of the key), so ranks are a permutation.  Only the coordinates
//     are written, in rank order; the key array's space then holds the neighbour grid.  The greedy min-distance pass and
//     the output are those of k_sort_select.  Lists this kernel cannot finish (overflow, too many candidates, a skewed
//     bin, or fewer than maxCorners corners among the kept candidates) are left to k_sort_select (done[list] stays 0).
// ---------------------------------------------------------------------------------------------
constexpr int FS_MAX_N = 16384, FS_MAX_BIN = 768;
constexpr int FS_MAX_CELLS = 32 * 1024;           // uint16 heads (candidate ranks are < FS_MAX_N): a finer grid than SS_MAX_CELLS allows
constexpr unsigned short FS_NIL = 0xffffu;
constexpr int TF_NT = 256, TF_PER_THREAD = 8, TF_CHUNK = TF_NT * TF_PER_THREAD, TF_CTAS = 32;

__device__ __forceinline__ int gftt_need(int stop_at) { return stop_at + stop_at / 2 + 256; }

// Smallest histogram bin `cut` with hist[0..cut] >= need (bins counted from the list's maximum downwards) and that count;
// NT threads, GF_BINS / NT consecutive bins per thread.  Optionally stores the exclusive prefix of every bin and the largest
// bin count among bins 0..cut.  s_out = {cut, n_top, max_bin}; cut = -1 if the whole histogram holds fewer than `need`.
template <int NT>
__device__ void find_cut(const int *__restrict__ hist, int need, int *s_scan, int *s_out, int *prefix)
{
    constexpr int PER = GF_BINS / NT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int h[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) { h[i] = hist[PER * tid + i]; sum += h[i]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (tid == 0) { s_out[0] = -1; s_out[1] = 0; s_out[2] = 0; }
    if (lane == 31) s_scan[wid] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; w++) woff += s_scan[w];
    int run = incl + woff - sum;                           // exclusive prefix of this thread's first bin
    int cut = -1, ntop = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        if (prefix) prefix[PER * tid + i] = run;
        if (run < need && need <= run + h[i]) { cut = PER * tid + i; ntop = run + h[i]; }
        run += h[i];
    }
    if (cut >= 0) { s_out[0] = cut; s_out[1] = ntop; }
    __syncthreads();
    {
        const int c = s_out[0];
        int mb = 0;
#pragma unroll
        for (int i = 0; i < PER; i++) if (c < 0 || PER * tid + i <= c) mb = max(mb, h[i]);
        mb = __reduce_max_sync(0xffffffffu, mb);
        if (lane == 0 && mb > 0) atomicMax(&s_out[2], mb);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TF_NT)
k_top_filter(const unsigned long long *__restrict__ keys_all, const int *__restrict__ counts, int cap, int max_corners, int max_out,
             const int32_t *__restrict__ enable, int n_masks, const unsigned *__restrict__ maxkey, const int *__restrict__ hist_all,
             unsigned long long *__restrict__ gtop_all, int *__restrict__ top_count)
{
    __shared__ int s_scan[TF_NT / 32], s_out[3], s_n, s_base;
    const int list = blockIdx.y, tid = threadIdx.x;
    if (enable && !enable[list / n_masks]) return;
    const int n_full = counts[list];
    if (n_full > cap) return;
    if (!(max_corners > 0 && max_corners <= max_out)) return;
    const int need = gftt_need(max_corners);
    if (!(n_full > need + need / 4)) return;
    find_cut<TF_NT>(hist_all + list * GF_BINS, need, s_scan, s_out, nullptr);
    const int cut = s_out[0];
    if (cut < 0) return;
    const unsigned kmax = maxkey[list];
    const unsigned long long *gk = keys_all + (long long)list * cap;
    unsigned long long *gt = gtop_all + (long long)list * cap;
    for (int blk = blockIdx.x; (long long)blk * TF_CHUNK < n_full; blk += gridDim.x) {
        if (tid == 0) s_n = 0;
        __syncthreads();
        unsigned long long k[TF_PER_THREAD];
        unsigned hit = 0;
#pragma unroll
        for (int j = 0; j < TF_PER_THREAD; j++) {
            const int i = blk * TF_CHUNK + j * TF_NT + tid;
            k[j] = i < n_full ? gk[i] : 0ull;
            if (i < n_full && key_bin(kmax, (unsigned)(k[j] >> 32)) <= cut) hit |= 1u << j;
        }
        int ofs = 0;
        if (hit) ofs = atomicAdd(&s_n, __popc(hit));
        __syncthreads();
        if (tid == 0 && s_n) s_base = atomicAdd(&top_count[list], s_n);
        __syncthreads();
        if (hit) {
            int pos = s_base + ofs;
#pragma unroll
            for (int j = 0; j < TF_PER_THREAD; j++)
                if ((hit >> j) & 1u) gt[pos++] = k[j];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SS_NT)
k_sort_select_fast(const unsigned long long *__restrict__ keys_all, const int *__restrict__ counts, int cap, int W, int H,
                   int max_corners, double min_distance, int cs, int gw, int gh, float *__restrict__ corners,
                   int32_t *__restrict__ out_counts, int max_out, const int32_t *__restrict__ enable, int n_masks,
                   const unsigned *__restrict__ maxkey, const int *__restrict__ hist_all,
                   const unsigned long long *__restrict__ gtop_all, const int *__restrict__ top_count, int *__restrict__ done)
{
    extern __shared__ __align__(16) unsigned char fs_smem[];
    __shared__ int s_scan[SS_NT / 32], s_out[3], s_base;
    unsigned long long *A = reinterpret_cast<unsigned long long *>(fs_smem);                 // FS_MAX_N keys (bin order)
    unsigned *xy = reinterpret_cast<unsigned *>(fs_smem + (size_t)FS_MAX_N * 8);             // FS_MAX_N coordinates (rank order)
    int *prefix = reinterpret_cast<int *>(fs_smem + (size_t)FS_MAX_N * 12);                  // GF_BINS
    int *fill = prefix + GF_BINS;                                                            // GF_BINS
    // after ranking: the neighbour grid lives in A's space, the states in prefix/fill's
    unsigned short *heads = reinterpret_cast<unsigned short *>(fs_smem);                     // <= FS_MAX_CELLS
    unsigned short *nxt = heads + FS_MAX_CELLS;                                              // FS_MAX_N
    volatile unsigned char *state = reinterpret_cast<volatile unsigned char *>(prefix);     // FS_MAX_N
    static_assert(FS_MAX_CELLS * 2 + FS_MAX_N * 2 <= FS_MAX_N * 8 && FS_MAX_N <= 2 * GF_BINS * 4 && FS_MAX_N < 0xffff, "overlay sizes");

    const int list = blockIdx.x, tid = threadIdx.x;
    if (enable && !enable[list / n_masks]) return;
    const int n_full = counts[list];
    if (n_full > cap) return;                                                                // overflow: general kernel
    const int stop_at = (max_corners > 0 && max_corners <= max_out) ? max_corners : 0x7fffffff;
    const int *hist = hist_all + list * GF_BINS;
    const unsigned kmax = maxkey ? maxkey[list] : 0u;
    bool trunc = false;
    int need = 0x7fffffff;
    if (stop_at != 0x7fffffff) {
        need = gftt_need(stop_at);
        trunc = n_full > need + need / 4;
    }
    find_cut<SS_NT>(hist, trunc ? need : 0x7fffffff, s_scan, s_out, prefix);
    fill[2 * tid] = 0; fill[2 * tid + 1] = 0;
    int n = n_full;
    const unsigned long long *src = keys_all + (long long)list * cap;
    if (trunc) {
        if (s_out[0] < 0) return;
        n = s_out[1];
        src = gtop_all + (long long)list * cap;
        if (top_count[list] != n) return;                                                    // (cannot happen; keeps the fallback safe)
    }
    if (n > FS_MAX_N || s_out[2] > FS_MAX_BIN) return;                                       // block-uniform
    __syncthreads();
    // counting sort by bin
    for (int i = tid; i < n; i += SS_NT) {
        const unsigned long long k = src[i];
        const int bin = key_bin(kmax, (unsigned)(k >> 32));
        A[prefix[bin] + atomicAdd(&fill[bin], 1)] = k;
    }
    __syncthreads();
    // exact rank inside the bin; coordinates in rank order
    for (int i = tid; i < n; i += SS_NT) {
        const unsigned long long k = A[i];
        const int bin = key_bin(kmax, (unsigned)(k >> 32));
        const int start = prefix[bin], end = start + fill[bin];
        int rank = start;
        for (int j = start; j < end; j++) rank += A[j] > k ? 1 : 0;
        const unsigned idx = (unsigned)k;
        const unsigned y = idx / (unsigned)W, x = idx - y * (unsigned)W;
        xy[rank] = (y << 16) | x;
    }
    __syncthreads();
    for (int i = tid; i < n; i += SS_NT) state[i] = (min_distance >= 1) ? 0 : 1;
    int accepted = n;
    if (min_distance >= 1) {
        for (int i = tid; i < (gw * gh + 1) / 2; i += SS_NT) reinterpret_cast<unsigned *>(heads)[i] = 0xffffffffu;
        __syncthreads();
        // Linked lists per cell, built by rank windows: inside a window the order of the 16-bit exchange emulation does not
        // matter, a cell's list is only ever traversed completely.
        for (int i = tid; i < n; i += SS_NT) {
            const unsigned v = xy[i];
            const int cell = ((v >> 16) / cs) * gw + (v & 0xffff) / cs;
            // atomic exchange on a 16-bit slot: CAS loop on the containing word
            unsigned *wp = reinterpret_cast<unsigned *>(heads) + (cell >> 1);
            const int sh = (cell & 1) * 16;
            unsigned old = *wp, assumed;
            do {
                assumed = old;
                old = atomicCAS(wp, assumed, (assumed & ~(0xffffu << sh)) | ((unsigned)i << sh));
            } while (old != assumed);
            nxt[i] = (unsigned short)((old >> sh) & 0xffffu);
        }
        __syncthreads();
        // integer coordinates: d2 is an exact integer, "d2 < minDistance^2" is "d2 < ceil(minDistance^2)"
        const double md2d = min_distance * min_distance;
        const int ithr = md2d >= 2.0e9 ? 0x7fffffff : (int)ceil(md2d);
        accepted = 0;
        for (int base = 0; base < n && accepted < stop_at; base += SS_NT) {
            const int i = base + tid;
            // higher-ranked candidates closer than minDistance: found once (first round), re-checked from registers
            int nb[4], n_nb = -1;
            for (;;) {
                int pending = 0;
                if (i < n && !state[i]) {
                    int verdict = 1;
                    if (n_nb < 0 || n_nb > 4) {
                        const unsigned v = xy[i];
                        const int x = v & 0xffff, y = v >> 16;
                        const int cx = x / cs, cy = y / cs;
                        int cntn = 0;
                        for (int yy = max(0, cy - 1); yy <= min(gh - 1, cy + 1); yy++)
                            for (int xx = max(0, cx - 1); xx <= min(gw - 1, cx + 1); xx++)
                                for (unsigned j = heads[yy * gw + xx]; j != FS_NIL; j = nxt[j]) {
                                    if ((int)j >= i) continue;
                                    const unsigned u = xy[j];
                                    const int dx = x - (int)(u & 0xffff), dy = y - (int)(u >> 16);
                                    if (dx * dx + dy * dy < ithr) {
                                        if (cntn < 4) nb[cntn] = (int)j;
                                        cntn++;
                                        const unsigned sj = state[j];
                                        if (sj == 1) verdict = 2;
                                        else if (sj == 0 && verdict != 2) verdict = 0;
                                    }
                                }
                        n_nb = cntn;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (q < n_nb) {
                                const unsigned sj = state[nb[q]];
                                if (sj == 1) verdict = 2;
                                else if (sj == 0 && verdict != 2) verdict = 0;
                            }
                    }
                    if (verdict) state[i] = (unsigned char)verdict; else pending = 1;
                }
                if (!__syncthreads_or(pending)) break;
            }
            accepted += __syncthreads_count(i < n && state[i] == 1);
        }
    }
    if (trunc && accepted < stop_at) return;                  // the kept candidates do not suffice: general kernel, full list
    __syncthreads();
    // rank-ordered compaction of the accepted candidates, cut at `limit`
    float *out = corners + (size_t)list * max_out * 2;
    const int limit = (max_corners > 0 && max_corners < max_out) ? max_corners : max_out;
    const bool may_truncate = max_corners <= 0 || max_corners > max_out;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SS_NT) {
        const int i = base + tid;
        const int acc = (i < n && state[i] == 1) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        const int lane = tid & 31, wid = tid >> 5;
        if (lane == 0) s_scan[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
        for (int w = 0; w < SS_NT / 32; w++) { const int c = s_scan[w]; if (w < wid) woff += c; total += c; }
        const int pos = s_base + woff + __popc(bal & ((1u << lane) - 1));
        if (acc && pos < limit) {
            const unsigned v = xy[i];
            out[2 * pos] = (float)(v & 0xffff); out[2 * pos + 1] = (float)(v >> 16);
        }
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
        if (s_base >= limit && !may_truncate) break;
    }
    if (tid == 0) {
        const int total = s_base;
        const int nc = total < limit ? total : limit;
        const bool truncated = may_truncate && total > limit;
        out_counts[list] = truncated ? -1 - nc : nc;
        done[list] = 1;
    }
}

}  // namespace

static int make_mask_set(const lkf_image *mask, int n_masks, int B, int H, int W, MaskSet *out)
{
    MaskSet ms{nullptr, 0, 0, 0};
    if (n_masks < 1 || n_masks > MAXM) return LKF_ERR_ARG;
    if (mask) {
        if (!img_ok(mask, 1, 1) || mask->height != H || mask->width != W) return LKF_ERR_ARG;
        if (mask->batch == n_masks) ms.per_image = 0;
        else if (mask->batch == B * n_masks) ms.per_image = 1;
        else return LKF_ERR_ARG;
        ms.p = (const uint8_t *)mask->data; ms.rp = mask->row_pitch; ms.ip = mask->image_pitch;
    } else if (n_masks != 1) return LKF_ERR_ARG;
    *out = ms;
    return LKF_OK;
}

// maxkey != NULL: also reduce the masked maxima (a6 step 1) while the map is produced; *fused tells whether that happened
static int min_eigen_impl(lkf_ctx *ctx, const lkf_image *src, const lkf_image *eig, int block_size, int dy_tail_cols, const int32_t *enable,
                          lkf_stream stream, MaskSet ms = MaskSet{nullptr, 0, 0, 0}, int n_masks = 1, unsigned *maxkey = nullptr,
                          bool *fused = nullptr)
{
    if (fused) *fused = false;
    if (!img_ok(src, 1, 1) || !img_ok(eig, 1, 4) || src->batch != eig->batch || src->height != eig->height ||
        src->width != eig->width || !aligned_to(eig, 4))
        return LKF_ERR_ARG;
    if (block_size < 1) return LKF_ERR_ARG;
    if (block_size > 31) return LKF_ERR_UNSUPPORTED;
    const int H = src->height, W = src->width, bs = block_size;
    if (dy_tail_cols < 0) dy_tail_cols = W % 32;
    const double scale = 1.0 / (4.0 * bs * 255.0);           // (1 << (ksize-1)) * blockSize * 255, ksize = 3
    const float s = (float)scale, s2 = (float)(2.0 * scale);
    if ((bs == 3 || bs == 5 || bs == 7) && W % 4 == 0 && W >= 8 && H >= 8 && aligned_to(src, 4)) {
        cudaStream_t st = (cudaStream_t)stream;
        int rc = bs == 3 ? launch_min_eigen_stream<3>(ctx, src, eig, dy_tail_cols, s, s2, enable, ms, n_masks, maxkey, st)
               : bs == 5 ? launch_min_eigen_stream<5>(ctx, src, eig, dy_tail_cols, s, s2, enable, ms, n_masks, maxkey, st)
                         : launch_min_eigen_stream<7>(ctx, src, eig, dy_tail_cols, s, s2, enable, ms, n_masks, maxkey, st);
        if (rc != LKF_OK) return rc;
        if (fused && maxkey) *fused = true;
        LKF_LAUNCH_CHECK(nullptr);
        return LKF_OK;
    }
    const int pw = ME_W + bs - 1, ph = ME_H + bs - 1, uw = pw + 2, uh = ph + 2, uws = (uw + 3) & ~3;
    size_t smem = (size_t)3 * ph * ME_W * 8 + (size_t)3 * ph * pw * 4 + (size_t)uh * uws;
    if (smem > 227 * 1024) return LKF_ERR_UNSUPPORTED;
    LKF_FUNC_ATTRS(LKF_CUDA(cudaFuncSetAttribute(k_min_eigen, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
    dim3 grid((W + ME_W - 1) / ME_W, (H + ME_H - 1) / ME_H, src->batch);
    k_min_eigen<<<grid, ME_NT, smem, (cudaStream_t)stream>>>((const uint8_t *)src->data, src->row_pitch, src->image_pitch, H, W,
                                                            (float *)eig->data, eig->row_pitch, eig->image_pitch, bs,
                                                            dy_tail_cols, s, s2, enable);
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}

LKF_EXPORT int lkf_corner_min_eigen_f32(const lkf_image *src, const lkf_image *eig, int block_size, int dy_tail_cols,
                                        lkf_stream stream)
{
    return min_eigen_impl(nullptr, src, eig, block_size, dy_tail_cols, nullptr, stream);
}

struct GfttScratch {
    unsigned long long *keys, *top; unsigned *maxkey; int *ccount, *hist, *topcount, *done; void *gaux;
    int cap, cs, gw, gh, lists;
    int fcs, fgw, fgh;          // neighbour grid of the fast selection kernel (finer: 16-bit heads)
};

// validates the selection arguments, sizes the scratch buffers and zeroes the per-list maxima / candidate counts
static int gftt_prepare(lkf_ctx *ctx, const lkf_gftt_params *p, int B, int H, int W, int n_masks, float *corners, int32_t *counts,
                        int max_out, lkf_stream stream, GfttScratch *g)
{
    if (!ctx || !p || !corners || !counts || max_out < 1) return LKF_ERR_ARG;
    if (!(p->quality_level > 0) || p->min_distance < 0) return LKF_ERR_ARG;          // cv2 asserts
    if (n_masks < 1 || n_masks > MAXM) return LKF_ERR_ARG;
    if (W > 65535 || H > 65535) return LKF_ERR_UNSUPPORTED;
    g->lists = B * n_masks;
    const long long P = (long long)H * W;
    long long cap_ll = P / 4 + 4096;
    g->cap = 32;
    while (g->cap < cap_ll) g->cap <<= 1;
    // neighbour grid: cell side >= minDistance (3x3 search suffices), grown until the heads fit in smem
    g->cs = 1; g->gw = 1; g->gh = 1;
    if (p->min_distance >= 1) {
        g->cs = (int)ceil(p->min_distance);
        for (;;) {
            g->gw = (W + g->cs - 1) / g->cs; g->gh = (H + g->cs - 1) / g->cs;
            if ((long long)g->gw * g->gh <= SS_MAX_CELLS) break;
            g->cs++;
        }
    }
    g->fcs = 1; g->fgw = 1; g->fgh = 1;
    if (p->min_distance >= 1) {
        g->fcs = (int)ceil(p->min_distance);
        for (;;) {
            g->fgw = (W + g->fcs - 1) / g->fcs; g->fgh = (H + g->fcs - 1) / g->fcs;
            if ((long long)g->fgw * g->fgh <= FS_MAX_CELLS) break;
            g->fcs++;
        }
    }
    void *keys, *cnt, *top;
    int rc;
    const size_t cnt_ints = (size_t)g->lists * (4 + GF_BINS);
    if ((rc = lkf_scratch(ctx, 0, (size_t)g->lists * g->cap * 8, &keys, (cudaStream_t)stream)) != LKF_OK) return rc;
    if ((rc = lkf_scratch(ctx, 1, cnt_ints * sizeof(int), &cnt, (cudaStream_t)stream)) != LKF_OK) return rc;
    if ((rc = lkf_scratch(ctx, 2, (size_t)g->lists * g->cap * 5, &g->gaux, (cudaStream_t)stream)) != LKF_OK) return rc;     // next (u32) + state (u8), large lists only
    if ((rc = lkf_scratch(ctx, 4, (size_t)g->lists * g->cap * 8, &top, (cudaStream_t)stream)) != LKF_OK) return rc;        // strongest candidates of large lists
    g->keys = (unsigned long long *)keys;
    g->top = (unsigned long long *)top;
    g->maxkey = (unsigned *)cnt;
    g->ccount = (int *)cnt + g->lists;
    g->topcount = (int *)cnt + 2 * g->lists;
    g->done = (int *)cnt + 3 * g->lists;
    g->hist = (int *)cnt + 4 * g->lists;
    LKF_CUDA(cudaMemsetAsync(cnt, 0, cnt_ints * sizeof(int), (cudaStream_t)stream));
    return LKF_OK;
}

static int gftt_passes(lkf_ctx *ctx, const lkf_gftt_params *p, const lkf_image *eig, MaskSet ms, int n_masks, const int32_t *enable,
                       float *corners, int32_t *counts, int max_out, lkf_stream stream, const GfttScratch &g, bool max_done)
{
    const int B = eig->batch, H = eig->height, W = eig->width;
    const long long P = (long long)H * W;
    cudaStream_t s = (cudaStream_t)stream;
    if (!max_done) {
        int chunks = (int)((P + 256 * 8 - 1) / (256 * 8));
        if (chunks > 4 * ctx->sm_count) chunks = 4 * ctx->sm_count;
        k_masked_max<<<dim3(chunks, B), 256, 0, s>>>((const float *)eig->data, eig->row_pitch, eig->image_pitch, H, W, ms, n_masks, g.maxkey, enable);
        LKF_LAUNCH_CHECK(ctx);
    }
    if (aligned_to(eig, 16)) {
        dim3 g4((W + 127) / 128, (H - 2 + 8 * NMS4_ROWS - 1) / (8 * NMS4_ROWS), B);
        k_nms_compact4<<<g4, 256, 0, s>>>((const float *)eig->data, eig->row_pitch, eig->image_pitch, H, W, ms, n_masks, g.maxkey,
                                          p->quality_level, g.keys, g.ccount, g.cap, enable, g.hist);
    } else {
        dim3 g2((W - 2 + 31) / 32, (H - 2 + 7) / 8, B);
        k_nms_compact<<<g2, 256, 0, s>>>((const float *)eig->data, eig->row_pitch, eig->image_pitch, H, W, ms, n_masks, g.maxkey,
                                         p->quality_level, g.keys, g.ccount, g.cap, enable, g.hist);
    }
    LKF_LAUNCH_CHECK(ctx);
    const size_t ss_smem = (size_t)SS_SMEM_KEYS * 8 + (size_t)SS_MAX_CELLS * 4 + (size_t)SS_SMEM_KEYS * 4 + SS_SMEM_KEYS;
    const size_t fs_smem = (size_t)FS_MAX_N * 12 + (size_t)GF_BINS * 8;
    LKF_FUNC_ATTRS({
        LKF_CUDA(cudaFuncSetAttribute(k_sort_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss_smem));
        LKF_CUDA(cudaFuncSetAttribute(k_sort_select_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs_smem));
    });
    k_top_filter<<<dim3(TF_CTAS, g.lists), TF_NT, 0, s>>>(g.keys, g.ccount, g.cap, p->max_corners, max_out, enable, n_masks, g.maxkey,
                                                         g.hist, g.top, g.topcount);
    LKF_LAUNCH_CHECK(ctx);
    k_sort_select_fast<<<g.lists, SS_NT, fs_smem, s>>>(g.keys, g.ccount, g.cap, W, H, p->max_corners, p->min_distance, g.fcs, g.fgw, g.fgh,
                                                      corners, counts, max_out, enable, n_masks, g.maxkey, g.hist, g.top, g.topcount,
                                                      g.done);
    LKF_LAUNCH_CHECK(ctx);
    // whatever the fast kernel left (done[list] == 0): capacity overflow, unlimited or very long lists, skewed bins
    k_sort_select<<<g.lists, SS_NT, ss_smem, s>>>(g.keys, g.ccount, g.cap, W, H, p->max_corners, p->min_distance, g.cs, g.gw,
                                                 g.gh, (unsigned *)g.gaux, (unsigned char *)g.gaux + (size_t)g.lists * g.cap * 4, corners,
                                                 counts, max_out, enable, n_masks, g.maxkey, g.hist, g.top, g.done);
    LKF_LAUNCH_CHECK(ctx);
    return LKF_OK;
}

static int gftt_select_impl(lkf_ctx *ctx, const lkf_gftt_params *p, const lkf_image *eig, const lkf_image *mask,
                            int n_masks, const int32_t *enable, float *corners, int32_t *counts, int max_out, lkf_stream stream)
{
    if (!img_ok(eig, 1, 4) || !aligned_to(eig, 4)) return LKF_ERR_ARG;
    const int B = eig->batch, H = eig->height, W = eig->width;
    MaskSet ms;
    int rc = make_mask_set(mask, n_masks, B, H, W, &ms);
    if (rc != LKF_OK) return rc;
    if (!ctx || !p || !corners || !counts || max_out < 1) return LKF_ERR_ARG;
    if (!(p->quality_level > 0) || p->min_distance < 0) return LKF_ERR_ARG;          // cv2 asserts
    if (H < 3 || W < 3) {        // no interior pixel: no corner
        LKF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * B * n_masks, (cudaStream_t)stream));
        return LKF_OK;
    }
    GfttScratch g;
    if ((rc = gftt_prepare(ctx, p, B, H, W, n_masks, corners, counts, max_out, stream, &g)) != LKF_OK) return rc;
    return gftt_passes(ctx, p, eig, ms, n_masks, enable, corners, counts, max_out, stream, g, false);
}

LKF_EXPORT int lkf_gftt_select(lkf_ctx *ctx, const lkf_gftt_params *p, const lkf_image *eig, const lkf_image *mask,
                               int n_masks, float *corners, int32_t *counts, int max_out, lkf_stream stream)
{
    return gftt_select_impl(ctx, p, eig, mask, n_masks, nullptr, corners, counts, max_out, stream);
}

// whole goodFeaturesToTrack for a batch: one eig map per image (workspace), n_masks selections per image,
// optional per-image predicate evaluated on the device (no host round trip for "refresh only where needed")
LKF_EXPORT int lkf_gftt_detect(lkf_ctx *ctx, const lkf_gftt_params *p, const lkf_image *src, const lkf_image *mask, int n_masks,
                               const int32_t *image_enable, float *corners, int32_t *counts, int max_out, lkf_stream stream)
{
    if (!ctx || !p || !img_ok(src, 1, 1)) return LKF_ERR_ARG;
    void *eigbuf;
    const size_t pitch = ((size_t)src->width * 4 + 15) & ~(size_t)15;
    int rc = lkf_scratch(ctx, 3, pitch * src->height * src->batch, &eigbuf, (cudaStream_t)stream);
    if (rc != LKF_OK) return rc;
    lkf_image eig = {eigbuf, src->batch, src->height, src->width, 1, (int64_t)pitch, (int64_t)(pitch * src->height)};
    const int B = src->batch, H = src->height, W = src->width;
    MaskSet ms;
    if ((rc = make_mask_set(mask, n_masks, B, H, W, &ms)) != LKF_OK) return rc;
    if (!corners || !counts || max_out < 1 || !(p->quality_level > 0) || p->min_distance < 0) return LKF_ERR_ARG;
    if (H < 3 || W < 3) {
        LKF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * B * n_masks, (cudaStream_t)stream));
        return LKF_OK;
    }
    GfttScratch g;
    if ((rc = gftt_prepare(ctx, p, B, H, W, n_masks, corners, counts, max_out, stream, &g)) != LKF_OK) return rc;
    bool fused = false;
    if ((rc = min_eigen_impl(ctx, src, &eig, p->block_size, p->dy_tail_cols, image_enable, stream, ms, n_masks, g.maxkey, &fused)) != LKF_OK)
        return rc;
    lkf_count_launch(ctx);
    return gftt_passes(ctx, p, &eig, ms, n_masks, image_enable, corners, counts, max_out, stream, g, fused);
}
