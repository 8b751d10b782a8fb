// lk_dense.cu -- a7 on the full integer grid (config C4: dense per-pixel flow, small odd window, 3-level pyramid):
// the same tracker as cv.calcOpticalFlowPyrLK (VP.py:545-546; oracle lko_pyrlk()) when prevPts is
// EVERY pixel (x, y) of the frame in raster order.  Results are bit-identical to k_lk_track_small on those points.
//
// What the grid buys (k_lk_track_small spends 14 000 thread-instruction slots per feature, 63 % of them active):
//   * At level l a grid feature sits at (x, y) * 2^-l: its window pixel (i, j) samples `prev` at
//     ((x + 2^l (i - R)) * 2^-l, (y + 2^l (j - R)) * 2^-l).  So ALL templates of a tile are views (stride 2^l) into one
//     "upsampled" map M_l(u, v) = Q14 bilinear sample of I and of its Scharr derivatives at (u, v) * 2^-l.  A CTA
//     builds M_l for its tile + halo once per level in shared memory (raw tile -> Scharr tile -> map), instead of every
//     thread re-deriving a 25-pixel template that overlaps its neighbour's by 24/25.
//   * No per-feature template state is left, so the Newton iterations can be re-packed: after every iteration the
//     still-active features of the tile are compacted (stable, raster order) and the next round runs them on the first
//     threads of the CTA -- converged features cost nothing, warps stay full (the thread-per-feature kernel keeps a warp
//     until its slowest feature has finished: 59 % lane utilisation of the iteration loops on the C4 frame).
//   * mismatch vector of an iteration = sum J~ (Ix~, Iy~) - sum I~ (Ix~, Iy~); the second term is a per-level constant of
//     the feature, so an iteration reads one packed gradient word per window pixel.
#include "lk_common.cuh"
#include <stdlib.h>

namespace {

constexpr int DN_TW = 32;      // tile width; tile height TH is a template parameter (features per CTA = threads per CTA = 32 TH)
constexpr int DN_MAX_LEVELS = 4;

template <int WIN, int DN_TH>
struct DN {
    static constexpr int R = (WIN - 1) / 2;
    static constexpr int halo(int l) { return R << l; }
    static constexpr int MW(int l) { return DN_TW + 2 * halo(l); }
    static constexpr int MH(int l) { return DN_TH + 2 * halo(l); }
    // integer sample grid under the map: (MW - 1) >> l spans + the bilinear neighbour + 1 for an unaligned origin
    static constexpr int IW(int l) { return ((MW(l) - 1) >> l) + 3; }
    static constexpr int IH(int l) { return ((MH(l) - 1) >> l) + 3; }
    static constexpr int RAWP(int l) { return (IW(l) + 2 + 3) & ~3; }
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int map_elems(int nl) { return nl <= 1 ? MW(0) * MH(0) : cmax(MW(nl - 1) * MH(nl - 1), map_elems(nl - 1)); }
    static constexpr int sch_elems(int nl) { return nl <= 1 ? IW(0) * IH(0) : cmax(IW(nl - 1) * IH(nl - 1), sch_elems(nl - 1)); }
    static constexpr int raw_bytes(int nl) { return nl <= 1 ? RAWP(0) * (IH(0) + 2) : cmax(RAWP(nl - 1) * (IH(nl - 1) + 2), raw_bytes(nl - 1)); }
};

// per-feature state of a tile (structure of arrays in shared memory)
template <int DN_NT>
struct DNState {
    float npx[DN_NT], npy[DN_NT];        // nextPts estimate at the current level (cv2's nextPts[ptidx] * 2 between levels)
    float nx[DN_NT], ny[DN_NT];          // iterate minus the half window
    float pdx[DN_NT], pdy[DN_NT];        // previous step (oscillation test)
    float A11[DN_NT], A12[DN_NT], A22[DN_NT], Dinv[DN_NT];
    int c1[DN_NT], c2[DN_NT];            // sum I~ Ix~, sum I~ Iy~ over the window (wrapping int32: see the kernel)
    float ev[DN_NT];
    unsigned short iters[DN_NT];
    unsigned char st[DN_NT], j[DN_NT];
    unsigned short list[2][DN_NT];
    int warp_cnt[DN_NT / 32];
    int n_next[3];
    int4 wlut[64];                       // Q14 weights of the (u mod 2^l, v mod 2^l) phases of the current level (<= 8 x 8)
};

// (WIN+1) x (WIN+1) bytes of a level at integer origin (ix, iy), REFLECT_101 in both directions, rows packed from byte 0:
// the border path of the window loader (a column's mirror index is found once, not once per row)
template <int WIN>
__device__ __noinline__ void load_window_edge(const uint8_t *__restrict__ P, long long rp, int W, int H, int ix, int iy, uint32_t *out)
{
    constexpr int NB = WIN + 1, NW = (NB + 3) / 4;
    int cx[NB];
#pragma unroll
    for (int c = 0; c < NB; c++) cx[c] = reflect101(ix + c, W);
#pragma unroll 1
    for (int y = 0; y < NB; y++) {
        const uint8_t *row = P + (long long)reflect101(iy + y, H) * rp;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            uint32_t v = 0;
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (4 * i + c < NB) v |= (uint32_t)row[cx[4 * i + c]] << (8 * c);
            out[y * NW + i] = v;
        }
    }
}

template <int WIN, int DN_TH>
size_t dn_smem_bytes(int n_levels)
{
    using D = DN<WIN, DN_TH>;
    size_t b = sizeof(DNState<DN_TW * DN_TH>);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)D::map_elems(n_levels) * 8;              // mapD (int2: Ix~, Iy~)
    b += ((size_t)D::map_elems(n_levels) * 2 + 15) & ~(size_t)15;   // mapI
    b += (size_t)D::sch_elems(n_levels) * 4;              // Scharr tile
    b += ((size_t)D::raw_bytes(n_levels) + 15) & ~(size_t)15;
    return b;
}


// Phase A of a level (kept OUT OF LINE: the kernel's iteration loop is compiled for 40 registers and its allocation is fragile):
// raw tile of prev -> Scharr tile -> the map M_l of the CTA's tile + halo.  NT threads, three barriers inside.
template <int WIN, int NT>
__device__ __noinline__ void dn_build_maps(const uint8_t *__restrict__ I, long long irp, int W, int H, int l, int x0, int y0, int th,
                                           unsigned char *raw, int *sch, int2 *mapD, unsigned short *mapI, int4 *wlut)
{
    constexpr int R = (WIN - 1) / 2;
    const int tid = threadIdx.x;
    const int Sd = 1 << l, halo = R << l;
    const int MW = DN_TW + 2 * halo, MH = th + 2 * halo;
    const int u0 = x0 - halo, v0 = y0 - halo;                  // map origin in level-0 ("upsampled") coordinates
    const int i0 = u0 >> l, j0 = v0 >> l;                      // integer sample under the map origin (floor)
    const int IW = ((u0 + MW - 1) >> l) - i0 + 2, IH = ((v0 + MH - 1) >> l) - j0 + 2;
    const int RWP = (IW + 2 + 3) & ~3;
    // (row, column) of a linear index < 65536 without a division: idx / d = (idx * ceil(2^32 / d)) >> 32 for d <= 256
    auto split = [](int idx, int d, unsigned magic, int &q, int &r) { q = (int)__umulhi((unsigned)idx, magic); r = idx - q * d; };
    // A1. raw tile of prev: rows j0-1 .. j0+IH, columns i0-1 .. i0+IW (REFLECT_101); the Q14 weights of the level's phases
    {
        const unsigned magic = 0xffffffffu / (unsigned)(IW + 2) + 1u;
        for (int idx = tid; idx < (IW + 2) * (IH + 2); idx += NT) {
            int j, i;
            split(idx, IW + 2, magic, j, i);
            raw[j * RWP + i] = I[(long long)reflect101(j0 - 1 + j, H) * irp + reflect101(i0 - 1 + i, W)];
        }
        if (tid < Sd * Sd) {
            const float sc = __int_as_float((127 - l) << 23);          // 2^-l, exact
            int w00, w01, w10, w11;
            weights(__fmul_rn((float)(tid & (Sd - 1)), sc), __fmul_rn((float)(tid >> l), sc), w00, w01, w10, w11);
            wlut[tid] = make_int4(w00, w01, w10, w11);
        }
    }
    __syncthreads();
    // A2. Scharr at the integer samples (i0 + a, j0 + c); ZERO outside the level (cv2's zero-bordered derivative image)
    {
        const unsigned magic = 0xffffffffu / (unsigned)IW + 1u;
        for (int idx = tid; idx < IW * IH; idx += NT) {
            int c, a;
            split(idx, IW, magic, c, a);
            const int gx = i0 + a, gy = j0 + c;
            int sx = 0, sy = 0;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
                const unsigned char *r0 = raw + c * RWP + a, *r1 = r0 + RWP, *r2 = r1 + RWP;
                sx = 3 * ((int)r0[2] - (int)r0[0]) + 10 * ((int)r1[2] - (int)r1[0]) + 3 * ((int)r2[2] - (int)r2[0]);
                sy = 3 * ((int)r2[0] - (int)r0[0]) + 10 * ((int)r2[1] - (int)r0[1]) + 3 * ((int)r2[2] - (int)r0[2]);
            }
            sch[idx] = (sx & 0xffff) | (sy << 16);
        }
    }
    __syncthreads();
    // A3. the map: Q14 bilinear samples of I (x32) and of the derivatives at (u, v) * 2^-l
    {
        const unsigned magic = 0xffffffffu / (unsigned)MW + 1u;
        for (int idx = tid; idx < MW * MH; idx += NT) {
            int mv, mu;
            split(idx, MW, magic, mv, mu);
            const int u = u0 + mu, v = v0 + mv;
            const int a = (u >> l) - i0, c = (v >> l) - j0;
            const int4 wq = wlut[((v & (Sd - 1)) << l) | (u & (Sd - 1))];
            const int w00 = wq.x, w01 = wq.y, w10 = wq.z, w11 = wq.w;
            const unsigned char *r0 = raw + (c + 1) * RWP + a + 1, *r1 = r0 + RWP;
            const int iv = descale((int)r0[0] * w00 + (int)r0[1] * w01 + (int)r1[0] * w10 + (int)r1[1] * w11, 9);
            const int d00 = sch[c * IW + a], d01 = sch[c * IW + a + 1], d10 = sch[(c + 1) * IW + a], d11 = sch[(c + 1) * IW + a + 1];
            const int ixv = descale((int)(short)d00 * w00 + (int)(short)d01 * w01 + (int)(short)d10 * w10 + (int)(short)d11 * w11, 14);
            const int iyv = descale((d00 >> 16) * w00 + (d01 >> 16) * w01 + (d10 >> 16) * w10 + (d11 >> 16) * w11, 14);
            mapI[idx] = (unsigned short)iv;
            mapD[idx] = make_int2(ixv, iyv);
        }
    }
    __syncthreads();
}

template <int WIN, int DN_TH, int MINB>
__global__ void __launch_bounds__(DN_TW * DN_TH, MINB)
k_lk_dense(LKLevels lv, int n_levels, int max_count, LKEps eps2, int flags, double min_eig_thr,
           float *__restrict__ next_pts, uint8_t *__restrict__ status, float *__restrict__ err,
           int *__restrict__ iters_out, unsigned al4_mask, int map_elems, int sch_elems,
           const int *__restrict__ grid_flag, int tail)
{
    using D = DN<WIN, DN_TH>;
    constexpr int DN_NT = DN_TW * DN_TH;
    constexpr int R = D::R, AREA = WIN * WIN, NWJ = (WIN + 1 + 3) / 4;
    static_assert(AREA <= 49, "exact int32 window sums");
    if (grid_flag && !*grid_flag) return;                 // prevPts is not the raster grid: the sparse kernel runs instead
    extern __shared__ __align__(16) unsigned char dn_smem[];
    DNState<DN_NT> &S = *reinterpret_cast<DNState<DN_NT> *>(dn_smem);
    int2 *mapD = reinterpret_cast<int2 *>(dn_smem + ((sizeof(DNState<DN_NT>) + 15) & ~(size_t)15));
    unsigned short *mapI = reinterpret_cast<unsigned short *>(mapD + map_elems);
    int *sch = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(mapI) + (((size_t)map_elems * 2 + 15) & ~(size_t)15));
    unsigned char *raw = reinterpret_cast<unsigned char *>(sch + sch_elems);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int W0 = lv.W[0], H0 = lv.H[0];
    const int x0 = blockIdx.x * DN_TW, y0 = blockIdx.y * DN_TH, b = blockIdx.z;
    const int tx = tid % DN_TW, ty = tid / DN_TW;
    const int fx = x0 + tx, fy = y0 + ty;
    const bool valid = fx < W0 && fy < H0;
    const float ppx = (float)fx, ppy = (float)fy;
    const float hx = (WIN - 1) * 0.5f, hy = (WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);

    S.npx[tid] = 0.f; S.npy[tid] = 0.f; S.ev[tid] = 0.f; S.iters[tid] = 0; S.st[tid] = 1;

    for (int l = n_levels - 1; l >= 0; --l) {
        const int W = lv.W[l], H = lv.H[l];
        const uint8_t *I = lv.prev[l] + (long long)b * lv.pip[l];
        const uint8_t *J = lv.next[l] + (long long)b * lv.nip[l];
        const long long irp = lv.prp[l], jrp = lv.nrp[l];
        const bool jal4 = (al4_mask >> l) & 1u;
        const int Sd = 1 << l, halo = R << l;
        const int MW = DN_TW + 2 * halo, MH = DN_TH + 2 * halo;
        __syncthreads();                                           // previous level's maps and lists are dead
        dn_build_maps<WIN, DN_NT>(I, irp, W, H, l, x0, y0, DN_TH, raw, sch, mapD, mapI, S.wlut);

        // B. level entry of this thread's feature: structure tensor, min-eigenvalue test, start iterate.
        //    Window pixel (i, j) of feature (tx, ty) is map sample (tx + i 2^l, ty + j 2^l).
        const float sc = __int_as_float((127 - l) << 23);
        bool active = false;
        if (valid) {
            float px = __fmul_rn(ppx, sc), py = __fmul_rn(ppy, sc);
            float nx, ny;
            if (l == n_levels - 1) { nx = px; ny = py; }
            else { nx = __fmul_rn(S.npx[tid], 2.f); ny = __fmul_rn(S.npy[tid], 2.f); }
            S.npx[tid] = nx; S.npy[tid] = ny;
            px = __fsub_rn(px, hx); py = __fsub_rn(py, hy);
            const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
            if (out_of_level(px, py, ipx, ipy, WIN, WIN, W, H)) {
                if (l == 0) { S.st[tid] = 0; S.ev[tid] = 0.f; }
            } else {
                int a11 = 0, a12 = 0, a22 = 0, c1 = 0, c2 = 0;
#pragma unroll
                for (int j = 0; j < WIN; j++) {
                    const int row = (ty + j * Sd) * MW + tx;
#pragma unroll
                    for (int i = 0; i < WIN; i++) {
                        const int2 d = mapD[row + i * Sd];
                        const int iv = mapI[row + i * Sd], ixv = d.x, iyv = d.y;
                        a11 += ixv * ixv; a12 += ixv * iyv; a22 += iyv * iyv;
                        c1 += iv * ixv; c2 += iv * iyv;
                    }
                }
                const float A11 = __fmul_rn((float)a11, FLT_SCALE), A12 = __fmul_rn((float)a12, FLT_SCALE), A22 = __fmul_rn((float)a22, FLT_SCALE);
                float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
                const float dd = __fsub_rn(A11, A22);
                const float disc = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
                const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * WIN * WIN));
                if (flags & LKF_OPTFLOW_LK_GET_MIN_EIGENVALS) S.ev[tid] = minEig;
                if ((double)minEig < min_eig_thr || Dt < 1.1920929e-07f) {
                    if (l == 0) S.st[tid] = 0;
                } else {
                    S.A11[tid] = A11; S.A12[tid] = A12; S.A22[tid] = A22; S.Dinv[tid] = __fdiv_rn(1.f, Dt);
                    S.c1[tid] = c1; S.c2[tid] = c2;
                    S.nx[tid] = __fsub_rn(nx, hx); S.ny[tid] = __fsub_rn(ny, hy);
                    S.pdx[tid] = 0.f; S.pdy[tid] = 0.f; S.j[tid] = 0;
                    active = max_count > 0;
                }
            }
        }
        // the first list: active features in raster order
        int cur = 0, n_act;
        {
            const unsigned bal = __ballot_sync(0xffffffffu, active);
            if (lane == 0) S.warp_cnt[wid] = __popc(bal);
            __syncthreads();
            int off = 0, total = 0;
#pragma unroll
            for (int w = 0; w < DN_NT / 32; w++) { const int c = S.warp_cnt[w]; if (w < wid) off += c; total += c; }
            if (active) S.list[0][off + __popc(bal & ((1u << lane) - 1))] = (unsigned short)tid;
            if (tid < 3) S.n_next[tid] = 0;
            n_act = total;
            __syncthreads();
        }

        // (WIN+1) rows of (WIN+1) bytes of `next` at integer origin (ix, iy), packed from byte 0 of each row's words
        auto load_window = [&](int ix, int iy, uint32_t (&jr)[WIN + 1][NWJ]) {
            if (jal4 && ix >= 0 && ix + WIN + 1 <= W) {
                // columns inside the level: aligned words (rows outside the level are mirrored).  The word holding the
                // last byte lies inside the 4-byte aligned row pitch; a further word is read only if a byte of it is used.
                const uint32_t *q0 = reinterpret_cast<const uint32_t *>(J + (long long)iy * jrp + ((ix >> 2) << 2));
                const int sh = (ix & 3) * 8, rw = (int)(jrp >> 2);
                const bool more = (ix & 3) + WIN + 1 > 4 * NWJ;
                if (iy >= 0 && iy + WIN < H) {
#pragma unroll
                    for (int y = 0; y <= WIN; y++) {
                        uint32_t t[NWJ + 1];
#pragma unroll
                        for (int i = 0; i < NWJ; i++) t[i] = __ldg(q0 + y * rw + i);
                        t[NWJ] = more ? __ldg(q0 + y * rw + NWJ) : 0u;
#pragma unroll
                        for (int i = 0; i < NWJ; i++) jr[y][i] = __funnelshift_r(t[i], t[i + 1], sh);
                    }
                } else {
#pragma unroll
                    for (int y = 0; y <= WIN; y++) {
                        const uint32_t *q = q0 + (reflect101(iy + y, H) - iy) * rw;
                        uint32_t t[NWJ + 1];
#pragma unroll
                        for (int i = 0; i < NWJ; i++) t[i] = __ldg(q + i);
                        t[NWJ] = more ? __ldg(q + NWJ) : 0u;
#pragma unroll
                        for (int i = 0; i < NWJ; i++) jr[y][i] = __funnelshift_r(t[i], t[i + 1], sh);
                    }
                }
            } else {
                uint32_t t[(WIN + 1) * NWJ];
                load_window_edge<WIN>(J, jrp, W, H, ix, iy, t);
#pragma unroll
                for (int y = 0; y <= WIN; y++)
#pragma unroll
                    for (int i = 0; i < NWJ; i++) jr[y][i] = t[y * NWJ + i];
            }
        };
        // one Newton iteration of feature f; true if the feature needs another one
        auto iterate = [&](int f) -> bool {
            const int ftx = f % DN_TW, fty = f / DN_TW;
            float nx = S.nx[f], ny = S.ny[f];
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (out_of_level(nx, ny, inx, iny, WIN, WIN, W, H)) {
                if (l == 0) S.st[f] = 0;
                return false;
            }
            int w00, w01, w10, w11;
            weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            const int wT = (w00 & 0xffff) | (w01 << 16), wB = (w10 & 0xffff) | (w11 << 16);
            // sum J~ Ix~, sum J~ Iy~ (each |.| <= 49 * 8160 * 4080 < 2^31); minus the constants in wrapping int32: the
            // true mismatch sums fit int32, so the wrapped difference is exact
            int b1 = 0, b2 = 0;
            uint32_t jr[WIN + 1][NWJ];
            load_window(inx, iny, jr);
            const int2 *mrow = mapD + fty * MW + ftx;
#pragma unroll
            for (int y = 0; y < WIN; y++) {
                int v[WIN];
                bilinear_packed<WIN, NWJ>(jr[y], jr[y + 1], 0, wT, wB, v);
#pragma unroll
                for (int x = 0; x < WIN; x++) {
                    const int2 d = mrow[x * Sd];
                    b1 += v[x] * d.x; b2 += v[x] * d.y;
                }
                mrow += Sd * MW;
            }
            b1 -= S.c1[f]; b2 -= S.c2[f];
            S.iters[f]++;
            const float A11 = S.A11[f], A12 = S.A12[f], A22 = S.A22[f], Dv = S.Dinv[f];
            const float fb1 = __fmul_rn((float)b1, FLT_SCALE), fb2 = __fmul_rn((float)b2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb2), __fmul_rn(A22, fb1)), Dv);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb1), __fmul_rn(A11, fb2)), Dv);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            float npx = __fadd_rn(nx, hx), npy = __fadd_rn(ny, hy);
            const int j = S.j[f];
            bool cont = false;
            if (lk_step_below_eps(dx, dy, eps2)) {
            } else if (j > 0 && lk_below_001(__fadd_rn(dx, S.pdx[f])) && lk_below_001(__fadd_rn(dy, S.pdy[f]))) {
                npx = __fsub_rn(npx, __fmul_rn(dx, 0.5f));
                npy = __fsub_rn(npy, __fmul_rn(dy, 0.5f));
            } else if (j + 1 < max_count) {
                cont = true;
                S.nx[f] = nx; S.ny[f] = ny; S.pdx[f] = dx; S.pdy[f] = dy; S.j[f] = (unsigned char)(j + 1);
            }
            S.npx[f] = npx; S.npy[f] = npy;
            return cont;
        };

        // C. Newton iterations, one per round; the unfinished features are re-packed after every round (one barrier per
        //    round: warp-aggregated append to the next list, three rotating counters).  Once `tail` or fewer are left, the
        //    warps holding them finish their features on their own (no more barriers).
        int round = 0;
        while (n_act > tail) {                                      // block-uniform
            bool cont = false;
            int f = 0;
            if (tid < n_act) { f = S.list[cur][tid]; cont = iterate(f); }
            const unsigned bal = __ballot_sync(0xffffffffu, cont);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&S.n_next[round % 3], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cont) S.list[cur ^ 1][base + __popc(bal & ((1u << lane) - 1))] = (unsigned short)f;
            if (tid == 0) S.n_next[(round + 1) % 3] = 0;            // the counter of the next round (last read two rounds ago)
            __syncthreads();
            n_act = S.n_next[round % 3];
            cur ^= 1; round++;
        }
        if (tid < n_act) {
            const int f = S.list[cur][tid];
            while (iterate(f)) { }
        }
        __syncthreads();

        // D. level 0: err = sum |J~ - I~| / (32 w h) at the final estimate
        if (l == 0 && !(flags & LKF_OPTFLOW_LK_GET_MIN_EIGENVALS) && valid && S.st[tid]) {
            const float ex = __fsub_rn(S.npx[tid], hx), ey = __fsub_rn(S.npy[tid], hy);
            const int iex = __float2int_rd(ex), iey = __float2int_rd(ey);
            if (out_of_level(ex, ey, iex, iey, WIN, WIN, W, H)) S.st[tid] = 0;
            else {
                int w00, w01, w10, w11;
                weights(__fsub_rn(ex, (float)iex), __fsub_rn(ey, (float)iey), w00, w01, w10, w11);
                const int wT = (w00 & 0xffff) | (w01 << 16), wB = (w10 & 0xffff) | (w11 << 16);
                int se = 0;
                uint32_t top[NWJ], bot[NWJ];
                load_row_bytes<WIN + 1>(J, jrp, W, H, jal4, iex, iey, top);
#pragma unroll
                for (int y = 0; y < WIN; y++) {
                    load_row_bytes<WIN + 1>(J, jrp, W, H, jal4, iex, iey + y + 1, bot);
                    int v[WIN];
                    bilinear_packed<WIN, NWJ>(top, bot, 0, wT, wB, v);
                    const int row = (ty + y) * MW + tx;          // l == 0: stride 1
#pragma unroll
                    for (int x = 0; x < WIN; x++) se += abs(v[x] - (int)mapI[row + x]);
#pragma unroll
                    for (int i = 0; i < NWJ; i++) top[i] = bot[i];
                }
                S.ev[tid] = __fdiv_rn((float)se, (float)(32 * WIN * WIN));
            }
        }
    }
    if (valid) {
        const long long pidx = (long long)b * W0 * H0 + (long long)fy * W0 + fx;
        const bool st = S.st[tid] != 0;
        reinterpret_cast<float2 *>(next_pts)[pidx] = make_float2(S.npx[tid], S.npy[tid]);
        status[pidx] = st ? 1 : 0;
        err[pidx] = st ? S.ev[tid] : 0.f;
        if (iters_out) iters_out[pidx] = S.iters[tid];
    }
}

// *flag = 1 iff prev_pts[b][i] == (i % W, i / W) for every point of every image (the raster grid of level 0)
__global__ void __launch_bounds__(256)
k_grid_check(const float2 *__restrict__ pts, long long n_total, int W, long long per_image, int *__restrict__ flag)
{
    bool ok = true;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
        const long long k = i % per_image;
        const int y = (int)(k / W), x = (int)(k - (long long)y * W);
        const float2 p = pts[i];
        ok &= p.x == (float)x && p.y == (float)y;
    }
    if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) *flag = 0;
}

template <int WIN, int DN_TH, int MINB>
int launch_dense(const LKLevels &lv, int n_levels, int max_count, double eps, const lkf_lk_params *p, float *next_pts, uint8_t *status,
                 float *err, int32_t *iters, unsigned al4_mask, int B, const int *grid_flag, cudaStream_t s)
{
    using D = DN<WIN, DN_TH>;
    constexpr int DN_NT = DN_TW * DN_TH;
    const size_t smem = dn_smem_bytes<WIN, DN_TH>(n_levels);
    if (smem > 227 * 1024) return LKF_ERR_UNSUPPORTED;
    LKF_FUNC_ATTRS({
        LKF_CUDA(cudaFuncSetAttribute(k_lk_dense<WIN, DN_TH, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dn_smem_bytes<WIN, DN_TH>(DN_MAX_LEVELS)));
        LKF_CUDA(cudaFuncSetAttribute(k_lk_dense<WIN, DN_TH, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    });
    // rounds are re-packed across the CTA while more than `tail` features are unfinished; then every warp finishes its own
    static const int tail = getenv("LKF_DN_TAIL") ? atoi(getenv("LKF_DN_TAIL")) : 32;
    dim3 grid((lv.W[0] + DN_TW - 1) / DN_TW, (lv.H[0] + DN_TH - 1) / DN_TH, B);
    k_lk_dense<WIN, DN_TH, MINB><<<grid, DN_NT, smem, s>>>(lv, n_levels, max_count, make_eps(eps), p->flags, p->min_eig_threshold, next_pts, status, err,
                                              iters, al4_mask, D::map_elems(n_levels), D::sch_elems(n_levels), grid_flag, tail);
    return LKF_OK;
}

}  // namespace

// true when the dense kernel covers this call: odd square window 3/5/7, at most DN_MAX_LEVELS levels, no initial flow
bool lk_dense_supported(const lkf_lk_params *p, int n_levels)
{
    return p->win_w == p->win_h && (p->win_w == 3 || p->win_w == 5 || p->win_w == 7) && n_levels >= 1 && n_levels <= DN_MAX_LEVELS &&
           !(p->flags & LKF_OPTFLOW_USE_INITIAL_FLOW);
}

// grid check of prev_pts into *flag (device); flag must hold 1 before (set here)
int lk_dense_grid_check(const float *prev_pts, int B, int W, int H, int *flag, cudaStream_t s)
{
    const long long per = (long long)W * H, n = per * B;
    LKF_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
    LKF_CUDA(cudaMemsetAsync(flag, 1, 1, s));                // little endian: the word now reads 1
    long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_grid_check<<<(int)blocks, 256, 0, s>>>(reinterpret_cast<const float2 *>(prev_pts), n, W, per, flag);
    return LKF_OK;
}

int lk_dense_launch(const lkf_lk_params *p, int n_levels, const lkf_image *prev_pyr, const lkf_image *next_pyr, int max_count, double eps,
                    float *next_pts, uint8_t *status, float *err, int32_t *iters, const int *grid_flag, cudaStream_t s)
{
    LKLevels lv;
    const int B = prev_pyr[0].batch;
    unsigned al4_mask = 0;
    for (int l = 0; l < n_levels; l++) {
        lv.prev[l] = (const uint8_t *)prev_pyr[l].data; lv.next[l] = (const uint8_t *)next_pyr[l].data;
        lv.prp[l] = prev_pyr[l].row_pitch; lv.pip[l] = prev_pyr[l].image_pitch;
        lv.nrp[l] = next_pyr[l].row_pitch; lv.nip[l] = next_pyr[l].image_pitch;
        lv.H[l] = prev_pyr[l].height; lv.W[l] = prev_pyr[l].width;
        if (aligned_to(&next_pyr[l], 4)) al4_mask |= 1u << l;
    }
    static const int minb = getenv("LKF_DN_MINB") ? atoi(getenv("LKF_DN_MINB")) : 6;      // resident CTAs per SM the registers are capped for (tuning switch)
#define LKF_DN_ARGS lv, n_levels, max_count, eps, p, next_pts, status, err, iters, al4_mask, B, grid_flag, s
    // measured on the C4 frame (1080p, 5x5): 0.903 / 0.848 / 0.847 ms for 4 / 5 / 6 resident CTAs (64 / 48 / 40 registers)
    if (p->win_w == 5 && minb == 4) return launch_dense<5, 8, 4>(LKF_DN_ARGS);
    switch (p->win_w) {
    case 3: return launch_dense<3, 8, 6>(LKF_DN_ARGS);
    case 5: return launch_dense<5, 8, 6>(LKF_DN_ARGS);
    case 7: return launch_dense<7, 8, 5>(LKF_DN_ARGS);
    }
#undef LKF_DN_ARGS
    return LKF_ERR_UNSUPPORTED;
}
