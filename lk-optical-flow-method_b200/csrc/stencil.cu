// stencil.cu -- a1 BGR->gray, a2 3x3 Gaussian, a1+a2 fused, a3 pyrDown, a4 Scharr (materialised).
// All integer, bit-exact against cv2 4.13 (formulas: SURVEY.md Appendix A.1-A.3; oracle: oracle/lkflow_oracle.c).
// Structure: shared-memory tiles with halos, 4-byte aligned bodies loaded as 32-bit words, REFLECT_101 by
// index mirroring (the mirror source of a 1..2-px halo is always inside the image).
#include "common.cuh"
#include "tma.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int NT = 256;

// ---------------------------------------------------------------------------------------------
// Tile loaders.  A tile row in shared memory holds WORDS 32-bit words; column c of the row is image
// column xs + c (xs is a multiple of 4).  Rows ys .. ys+ROWS-1 (mirrored).
// ---------------------------------------------------------------------------------------------
struct SrcU8 {
    const uint8_t *p; long long rp; int H, W; bool al4;
    __device__ __forceinline__ uint32_t word(int y, int gx) const
    {
        const uint8_t *row = p + (long long)y * rp;
        if (al4 && gx >= 0 && gx + 3 < W) return *reinterpret_cast<const uint32_t *>(row + gx);
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v |= (uint32_t)row[reflect101(gx + j, W)] << (8 * j);
        return v;
    }
};
// gray computed on the fly from interleaved BGR (fused a1+a2)
struct SrcBGR {
    const uint8_t *p; long long rp; int H, W; bool al4;
    __device__ __forceinline__ uint32_t word(int y, int gx) const
    {
        const uint8_t *row = p + (long long)y * rp;
        if (al4 && gx >= 0 && gx + 3 < W) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(row + 3 * gx);
            uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
            // bytes: a = B0 G0 R0 B1 | b = G1 R1 B2 G2 | c = R2 B3 G3 R3
            uint32_t g0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
            uint32_t g1 = gray_of(a >> 24, b & 255, (b >> 8) & 255);
            uint32_t g2 = gray_of((b >> 16) & 255, b >> 24, c & 255);
            uint32_t g3 = gray_of((c >> 8) & 255, (c >> 16) & 255, c >> 24);
            return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
        }
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint8_t *px = row + 3 * reflect101(gx + j, W);
            v |= gray_of(px[0], px[1], px[2]) << (8 * j);
        }
        return v;
    }
};

template <int ROWS, int WORDS, class Src>
__device__ __forceinline__ void load_tile(const Src &s, int xs, int ys, uint32_t (*sm)[WORDS])
{
    for (int i = threadIdx.x; i < ROWS * WORDS; i += NT) {
        int r = i / WORDS, k = i - r * WORDS;
        sm[r][k] = s.word(reflect101(ys + r, s.H), xs + 4 * k);
    }
}

__device__ __forceinline__ void store4(uint8_t *row, int x, int W, bool al4, uint32_t v)
{
    if (al4 && x + 3 < W) *reinterpret_cast<uint32_t *>(row + x) = v;
    else
        for (int j = 0; j < 4 && x + j < W; j++) row[x + j] = (uint8_t)(v >> (8 * j));
}

// ---------------------------------------------------------------------------------------------
// a1: BGR -> gray, 16 px (48 B in, 16 B out) per thread when 16-byte aligned, else 4 px / scalar.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_bgr2gray(Img src, Img dst, int al16, int al4)
{
    const int y = blockIdx.y, b = blockIdx.z;
    const uint8_t *srow = src.p + (long long)b * src.ip + (long long)y * src.rp;
    uint8_t *drow = dst.p + (long long)b * dst.ip + (long long)y * dst.rp;
    const int W = src.W;
    int x = (blockIdx.x * NT + threadIdx.x) * 16;
    if (x >= W) return;
    if (al16 && x + 16 <= W) {
        const uint4 *q = reinterpret_cast<const uint4 *>(srow + 3 * x);
        uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
        uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t a = w[3 * k], bb = w[3 * k + 1], c = w[3 * k + 2];
            uint32_t g0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
            uint32_t g1 = gray_of(a >> 24, bb & 255, (bb >> 8) & 255);
            uint32_t g2 = gray_of((bb >> 16) & 255, bb >> 24, c & 255);
            uint32_t g3 = gray_of((c >> 8) & 255, (c >> 16) & 255, c >> 24);
            o[k] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
        }
        *reinterpret_cast<uint4 *>(drow + x) = make_uint4(o[0], o[1], o[2], o[3]);
        return;
    }
    SrcBGR s{srow, 0, 1, W, al4 != 0};
    for (int k = 0; k < 4 && x + 4 * k < W; k++) store4(drow, x + 4 * k, W, al4 != 0, s.word(0, x + 4 * k));
}

// ---------------------------------------------------------------------------------------------
// a2 / a1+a2: 3x3 [1 2 1]^2, (sum+8)>>4, tile 128 x 32 out, 1-px halo.  Each thread: 4 px x 4 rows.
// ---------------------------------------------------------------------------------------------
constexpr int BT_W = 128, BT_H = 32;
constexpr int BT_WORDS = BT_W / 4 + 2;          // one pad word each side (halo lives in it)

template <class Src>
__global__ void __launch_bounds__(NT) k_blur3(Src src0, long long src_ip, Img dst, int dst_al4)
{
    __shared__ uint32_t sm[BT_H + 2][BT_WORDS];
    const int x0 = blockIdx.x * BT_W, y0 = blockIdx.y * BT_H, b = blockIdx.z;
    Src s = src0;
    s.p += (long long)b * src_ip;
    load_tile<BT_H + 2, BT_WORDS>(s, x0 - 4, y0 - 1, sm);
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8 threads, 4 rows each
    const int x = x0 + 4 * tx;
    if (x >= dst.W) return;
    // horizontal sums h[r][j] = p[x+j-1] + 2 p[x+j] + p[x+j+1] for the 6 rows feeding 4 output rows
    uint32_t h[6][4];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        const uint32_t *row = sm[4 * ty + r];
        uint32_t l = row[tx], m = row[tx + 1], rr = row[tx + 2];
        uint32_t p[6] = {l >> 24, m & 255, (m >> 8) & 255, (m >> 16) & 255, m >> 24, rr & 255};
#pragma unroll
        for (int j = 0; j < 4; j++) h[r][j] = p[j] + 2 * p[j + 1] + p[j + 2];
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int y = y0 + 4 * ty + r;
        if (y >= dst.H) break;
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v |= ((h[r][j] + 2 * h[r + 1][j] + h[r + 2][j] + 8) >> 4) << (8 * j);
        store4(dst.p + (long long)b * dst.ip + (long long)y * dst.rp, x, dst.W, dst_al4 != 0, v);
    }
}

// ---------------------------------------------------------------------------------------------
// a3: pyrDown 5x5 [1 4 6 4 1]^2, (sum+128)>>8, even source coordinates.  Output tile 64 x 16.
// ---------------------------------------------------------------------------------------------
constexpr int PD_W = 64, PD_H = 16;
constexpr int PD_SROWS = 2 * PD_H + 3;           // source rows 2*y0-2 .. 2*(y0+PD_H-1)+2
constexpr int PD_SWORDS = (2 * PD_W + 8) / 4;    // source cols 2*x0-4 .. 2*x0+2*PD_W+3

__global__ void __launch_bounds__(NT) k_pyrdown(SrcU8 src0, long long src_ip, Img dst, int dst_al4)
{
    __shared__ uint32_t sm[PD_SROWS][PD_SWORDS];
    __shared__ uint16_t hs[PD_SROWS][PD_W];
    const int x0 = blockIdx.x * PD_W, y0 = blockIdx.y * PD_H, b = blockIdx.z;
    SrcU8 s = src0;
    s.p += (long long)b * src_ip;
    load_tile<PD_SROWS, PD_SWORDS>(s, 2 * x0 - 4, 2 * y0 - 2, sm);
    __syncthreads();
    // horizontal pass: out col xo reads source cols 2xo-2..2xo+2 = tile cols 2(xo-x0)+2 .. +6
    for (int i = threadIdx.x; i < PD_SROWS * (PD_W / 2); i += NT) {
        int r = i / (PD_W / 2), k = i - r * (PD_W / 2);     // k: pair of outputs (2k, 2k+1)
        const uint32_t *row = sm[r];
        uint32_t a = row[k], m = row[k + 1], c = row[k + 2];
        // tile cols 4k .. 4k+11 ; outputs 2k -> cols 4k+2..4k+6 ; 2k+1 -> cols 4k+4..4k+8
        uint32_t p2 = (a >> 16) & 255, p3 = a >> 24, p4 = m & 255, p5 = (m >> 8) & 255, p6 = (m >> 16) & 255,
                 p7 = m >> 24, p8 = c & 255;
        hs[r][2 * k] = (uint16_t)(p2 + 4 * p3 + 6 * p4 + 4 * p5 + p6);
        hs[r][2 * k + 1] = (uint16_t)(p4 + 4 * p5 + 6 * p6 + 4 * p7 + p8);
    }
    __syncthreads();
    // vertical pass: 4 consecutive outputs per thread
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16
    const int x = x0 + 4 * tx, y = y0 + ty;
    if (x >= dst.W || y >= dst.H) return;
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int c = 4 * tx + j;
        uint32_t sacc = hs[2 * ty][c] + 4u * hs[2 * ty + 1][c] + 6u * hs[2 * ty + 2][c] + 4u * hs[2 * ty + 3][c] +
                        hs[2 * ty + 4][c];
        v |= ((sacc + 128) >> 8) << (8 * j);
    }
    store4(dst.p + (long long)b * dst.ip + (long long)y * dst.rp, x, dst.W, dst_al4 != 0, v);
}

// ---------------------------------------------------------------------------------------------
// a4: Scharr, int16 (Ix,Iy) interleaved, unnormalised; tile 128 x 32, 1-px halo (as the blur).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_scharr(SrcU8 src0, long long src_ip, Img dst, int dst_al16)
{
    __shared__ uint32_t sm[BT_H + 2][BT_WORDS];
    const int x0 = blockIdx.x * BT_W, y0 = blockIdx.y * BT_H, b = blockIdx.z;
    SrcU8 s = src0;
    s.p += (long long)b * src_ip;
    load_tile<BT_H + 2, BT_WORDS>(s, x0 - 4, y0 - 1, sm);
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = x0 + 4 * tx;
    if (x >= dst.W) return;
    int p[6][6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        const uint32_t *row = sm[4 * ty + r];
        uint32_t l = row[tx], m = row[tx + 1], rr = row[tx + 2];
        p[r][0] = l >> 24; p[r][1] = m & 255; p[r][2] = (m >> 8) & 255; p[r][3] = (m >> 16) & 255;
        p[r][4] = m >> 24; p[r][5] = rr & 255;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int y = y0 + 4 * ty + r;
        if (y >= dst.H) break;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int sp = 3 * p[r][j + 2] + 10 * p[r + 1][j + 2] + 3 * p[r + 2][j + 2];
            int sn = 3 * p[r][j] + 10 * p[r + 1][j] + 3 * p[r + 2][j];
            int ix = sp - sn;
            int iy = 3 * (p[r + 2][j] - p[r][j]) + 10 * (p[r + 2][j + 1] - p[r][j + 1]) + 3 * (p[r + 2][j + 2] - p[r][j + 2]);
            o[j] = ((uint32_t)ix & 0xffffu) | ((uint32_t)iy << 16);
        }
        uint8_t *drow = dst.p + (long long)b * dst.ip + (long long)y * dst.rp;
        if (dst_al16 && x + 3 < dst.W) *reinterpret_cast<uint4 *>(drow + 4 * x) = make_uint4(o[0], o[1], o[2], o[3]);
        else
            for (int j = 0; j < 4 && x + j < dst.W; j++) *reinterpret_cast<uint32_t *>(drow + 4 * (x + j)) = o[j];
    }
}

// ---------------------------------------------------------------------------------------------
// a1+a2 (and a2) streaming form for 16-byte aligned rows with W % 16 == 0 (1080p, 4K, pyramid levels):
// no shared memory, no barriers.  A thread owns 16 consecutive columns and slides down R rows; per input
// row it loads 48 B of BGR as 3 x 128-bit (or 16 B of gray), converts 16 + 2 pixels (the two neighbours
// come from one aligned 32-bit load each), forms the horizontal [1 2 1] sums packed two per register
// and keeps the last two rows of sums in registers; the vertical [1 2 1] pass, rounding and byte packing
// run on the packed pairs.  Gray uses two DP4A with the weights split as w = 128*wh + wl.
// ---------------------------------------------------------------------------------------------
constexpr int SV_NT = 128, SV_PX = 16;

// Gray of interleaved BGR with DP2A (two 16-bit weights x two bytes per instruction): the four pixels of three words
// a = B0 G0 R0 B1 | b = G1 R1 B2 G2 | c = R2 B3 G3 R3 need no byte shuffling at all -- every pixel is one .lo and one .hi
// dot product with one of four constant weight pairs.  Y = (3735 B + 19235 G + 9798 R + 2^14) >> 15.
constexpr uint32_t GW_BG = 3735u | (19235u << 16), GW_R_ = 9798u, GW__B = 3735u << 16, GW_GR = 19235u | (9798u << 16);
__device__ __forceinline__ uint32_t gray_px0(uint32_t a) { return __dp2a_hi(GW_R_, a, __dp2a_lo(GW_BG, a, 16384u)) >> 15; }             // bytes 0,1,2 of a
__device__ __forceinline__ uint32_t gray_px1(uint32_t a, uint32_t b) { return __dp2a_lo(GW_GR, b, __dp2a_hi(GW__B, a, 16384u)) >> 15; } // a3 b0 b1
__device__ __forceinline__ uint32_t gray_px2(uint32_t b, uint32_t c) { return __dp2a_lo(GW_R_, c, __dp2a_hi(GW_BG, b, 16384u)) >> 15; } // b2 b3 c0
__device__ __forceinline__ uint32_t gray_px3(uint32_t c) { return __dp2a_hi(GW_GR, c, __dp2a_lo(GW__B, c, 16384u)) >> 15; }             // bytes 1,2,3 of c

// 16 gray values of columns x0..x0+15 plus the neighbours g[-1], g[16]; g as 18 ints (index j+1)
struct RowBGR {
    __device__ __forceinline__ static void load(const uint8_t *__restrict__ row, int x0, int W, uint32_t (&g)[18])
    {
        const uint4 *q = reinterpret_cast<const uint4 *>(row + 3 * x0);
        const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
        const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t a = w[3 * k], b = w[3 * k + 1], c = w[3 * k + 2];
            g[4 * k + 1] = gray_px0(a);
            g[4 * k + 2] = gray_px1(a, b);
            g[4 * k + 3] = gray_px2(b, c);
            g[4 * k + 4] = gray_px3(c);
        }
        if (x0 > 0) g[0] = gray_px3(__ldg(reinterpret_cast<const uint32_t *>(row + 3 * x0 - 4)));
        else g[0] = g[2];                                                // REFLECT_101: column -1 -> 1
        if (x0 + SV_PX < W) g[17] = gray_px0(__ldg(reinterpret_cast<const uint32_t *>(row + 3 * x0 + 48)));
        else g[17] = g[15];                                              // column W -> W-2
    }
};
struct RowU8 {
    __device__ __forceinline__ static void load(const uint8_t *__restrict__ row, int x0, int W, uint32_t (&g)[18])
    {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row + x0));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            g[4 * k + 1] = w[k] & 255u; g[4 * k + 2] = (w[k] >> 8) & 255u; g[4 * k + 3] = (w[k] >> 16) & 255u; g[4 * k + 4] = w[k] >> 24;
        }
        g[0] = x0 > 0 ? (uint32_t)__ldg(row + x0 - 1) : g[2];
        g[17] = x0 + SV_PX < W ? (uint32_t)__ldg(row + x0 + 16) : g[15];
    }
};

template <class Row>
__device__ __forceinline__ void hsum_row(const uint8_t *__restrict__ row, int x0, int W, uint32_t (&h)[8])
{
    uint32_t g[18];
    Row::load(row, x0, W, g);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t a = g[2 * k] + 2u * g[2 * k + 1] + g[2 * k + 2];       // column 2k
        uint32_t b = g[2 * k + 1] + 2u * g[2 * k + 2] + g[2 * k + 3];   // column 2k+1
        h[k] = a | (b << 16);
    }
}

template <class Row>
__global__ void __launch_bounds__(SV_NT) k_blur3_stream(const uint8_t *__restrict__ src, long long srp, long long sip,
                                                        uint8_t *__restrict__ dst, long long drp, long long dip, int H, int W,
                                                        int rows_per_block)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * SV_PX;
    if (x0 >= W) return;
    const int b = blockIdx.z;
    const uint8_t *S = src + (long long)b * sip;
    uint8_t *D = dst + (long long)b * dip + x0;
    const int y0 = blockIdx.y * rows_per_block, y1 = min(H, y0 + rows_per_block);
    uint32_t hm[8], hc[8], hp[8];
    hsum_row<Row>(S + (long long)reflect101(y0 - 1, H) * srp, x0, W, hm);
    hsum_row<Row>(S + (long long)y0 * srp, x0, W, hc);
    for (int y = y0; y < y1; y++) {
        hsum_row<Row>(S + (long long)reflect101(y + 1, H) * srp, x0, W, hp);
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // packed pairs: (hm + 2 hc + hp + 8) >> 4, no carry between the 16-bit halves (max 4080 + 8)
            uint32_t v0 = ((hm[2 * k] + hp[2 * k] + 0x00080008u + 2u * hc[2 * k]) >> 4) & 0x00ff00ffu;
            uint32_t v1 = ((hm[2 * k + 1] + hp[2 * k + 1] + 0x00080008u + 2u * hc[2 * k + 1]) >> 4) & 0x00ff00ffu;
            o[k] = __byte_perm(v0, v1, 0x6420);
        }
        *reinterpret_cast<uint4 *>(D + (long long)y * drp) = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int k = 0; k < 8; k++) { hm[k] = hc[k]; hc[k] = hp[k]; }
    }
}

// threads per CTA of the streaming kernels: the columns of one image row rounded up to whole warps, at most SV_NT
static inline int stream_threads(int columns_of_threads)
{
    int nt = (columns_of_threads + 31) & ~31;
    return nt < 32 ? 32 : (nt > SV_NT ? SV_NT : nt);
}

template <class Row>
int launch_blur3_stream(const lkf_image *src, const lkf_image *dst, cudaStream_t s)
{
    const int H = dst->height, W = dst->width, R = 32;
    const int nt = stream_threads(W / SV_PX);                  // narrow images: smaller CTAs, no idle warps
    dim3 grid((W / SV_PX + nt - 1) / nt, (H + R - 1) / R, dst->batch);
    k_blur3_stream<Row><<<grid, nt, 0, s>>>((const uint8_t *)src->data, src->row_pitch, src->image_pitch, (uint8_t *)dst->data,
                                               dst->row_pitch, dst->image_pitch, H, W, R);
    return LKF_OK;
}

bool stream_ok(const lkf_image *src, const lkf_image *dst)
{
    return dst->width % SV_PX == 0 && dst->width >= 2 * SV_PX && dst->height >= 2 && aligned_to(src, 16) && aligned_to(dst, 16);
}

// ---------------------------------------------------------------------------------------------
// a3 streaming form (source rows 16-byte aligned, W_src % 16 == 0): a thread owns 8 output columns
// (16 source columns + 2 left + 1 right neighbour) and slides down, two output rows per iteration.
// Horizontal [1 4 6 4 1] at the even centres by DP4A on the raw bytes: the four bytes c-2..c+1 are a
// source word (odd outputs) or two words funnel-shifted by 16 bits (even outputs), dotted with the
// packed weights (1,4,6,4); the fifth tap s[c+2] is a second DP4A with a one-hot weight word.  The
// sums (<= 4080) are packed two per register, five rows of them live in registers, and the vertical
// pass + rounding run on the packed pairs (16*4080 + 128 < 2^16, no carry between the halves).
// ---------------------------------------------------------------------------------------------
// raw bytes of one source row for a thread: 16 columns, the word left of them (its bytes 2, 3 = columns -2, -1) and the
// byte right of them; REFLECT_101 at the left / right image border is resolved here
struct PdRaw { uint4 v; uint32_t l, r; };
__device__ __forceinline__ PdRaw pd_load_row(const uint8_t *__restrict__ S, long long srp, int y, int Hs, int xs, int Ws)
{
    if (y < 0) y = -y; else if (y >= Hs) y = 2 * (Hs - 1) - y;            // REFLECT_101, |overshoot| <= 2 < Hs
    const uint8_t *row = S + (long long)y * srp;
    PdRaw q;
    q.v = __ldg(reinterpret_cast<const uint4 *>(row + xs));
    q.l = xs > 0 ? __ldg(reinterpret_cast<const uint32_t *>(row + xs - 4)) : __byte_perm(q.v.x, 0u, 0x1200);   // columns -2, -1 -> 2, 1
    q.r = xs + 16 < Ws ? (uint32_t)__ldg(row + xs + 16) : (q.v.w >> 16) & 255u;                               // column Ws -> Ws-2
    return q;
}

// same, for a row known to be inside the image: p = address of column xs of the row (warp-uniform fast path of the loop)
__device__ __forceinline__ PdRaw pd_load_row_at(const uint8_t *__restrict__ p, bool has_left, bool has_right)
{
    PdRaw q;
    q.v = __ldg(reinterpret_cast<const uint4 *>(p));
    q.l = 0u; q.r = 0u;
    if (has_left) q.l = __ldg(reinterpret_cast<const uint32_t *>(p - 4));
    if (has_right) q.r = (uint32_t)__ldg(p + 16);
    if (!has_left) q.l = __byte_perm(q.v.x, 0u, 0x1200);
    if (!has_right) q.r = (q.v.w >> 16) & 255u;
    return q;
}

__device__ __forceinline__ void pd_hsum_row(const PdRaw &q, uint32_t (&h)[4])
{
    const uint32_t w[5] = {q.v.x, q.v.y, q.v.z, q.v.w, q.r};
    uint32_t left = q.l;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        // outputs 2k (centre = byte 0 of w[k]) and 2k+1 (centre = byte 2 of w[k])
        const uint32_t f = __funnelshift_r(left, w[k], 16);                       // bytes c-2 .. c+1 of the even output
        const uint32_t e = __dp4a(f, 0x04060401u, __dp4a(w[k], 0x00010000u, 0u));
        const uint32_t o = __dp4a(w[k], 0x04060401u, __dp4a(w[k + 1], 0x00000001u, 0u));
        h[k] = o * 65536u + e;
        left = w[k];
    }
}

__global__ void __launch_bounds__(SV_NT, 8) k_pyrdown_stream(const uint8_t *__restrict__ src, long long srp, long long sip,
                                                             uint8_t *__restrict__ dst, long long drp, long long dip, int Hs, int Ws,
                                                             int Hd, int Wd, int rows_per_block)
{
    const int xo = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (xo >= Wd) return;
    const int b = blockIdx.z, xs = 2 * xo;
    const uint8_t *S = src + (long long)b * sip;
    uint8_t *D = dst + (long long)b * dip + xo;
    const int y0 = blockIdx.y * rows_per_block, y1 = min(Hd, y0 + rows_per_block);
    const bool has_left = xs > 0, has_right = xs + 16 < Ws;
    uint32_t r0[4], r1[4], r2[4], r3[4], r4[4], r5[4], r6[4];
    {
        const PdRaw q0 = pd_load_row(S, srp, 2 * y0 - 2, Hs, xs, Ws), q1 = pd_load_row(S, srp, 2 * y0 - 1, Hs, xs, Ws),
                    q2 = pd_load_row(S, srp, 2 * y0, Hs, xs, Ws);
        pd_hsum_row(q0, r0); pd_hsum_row(q1, r1); pd_hsum_row(q2, r2);
    }
    auto vstore = [&](uint8_t *drow, const uint32_t (&a)[4], const uint32_t (&bb)[4], const uint32_t (&c)[4], const uint32_t (&d)[4],
                      const uint32_t (&e)[4]) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            o[k] = ((a[k] + e[k] + 0x00800080u + 4u * (bb[k] + d[k]) + 6u * c[k]) >> 8) & 0x00ff00ffu;
        *reinterpret_cast<uint2 *>(drow) = make_uint2(__byte_perm(o[0], o[1], 0x6420), __byte_perm(o[2], o[3], 0x6420));
    };
    // Two output rows per iteration; the two source rows of an output row are loaded while the previous output row is
    // computed.  Rows 2y+1 .. 2y+6 of an iteration are inside the image except in the last block(s) of an image
    // (warp-uniform): pointer bumps on the fast path, mirrored row indices otherwise (an overshoot of up to 4 rows past
    // the rows actually used still mirrors into the image: Hs >= 4).
    const uint8_t *p = S + (long long)(2 * y0 + 1) * srp + xs;            // row 2y+1 of the current iteration
    uint8_t *drow = D + (long long)y0 * drp;
    PdRaw qa = pd_load_row(S, srp, 2 * y0 + 1, Hs, xs, Ws), qb = pd_load_row(S, srp, 2 * y0 + 2, Hs, xs, Ws);
    for (int y = y0; y < y1; y += 2) {
        const bool fast = 2 * y + 6 < Hs;
        PdRaw qc, qd;
        if (fast) { qc = pd_load_row_at(p + 2 * srp, has_left, has_right); qd = pd_load_row_at(p + 3 * srp, has_left, has_right); }
        else { qc = pd_load_row(S, srp, 2 * y + 3, Hs, xs, Ws); qd = pd_load_row(S, srp, 2 * y + 4, Hs, xs, Ws); }
        pd_hsum_row(qa, r3); pd_hsum_row(qb, r4);
        vstore(drow, r0, r1, r2, r3, r4);
        if (fast) { qa = pd_load_row_at(p + 4 * srp, has_left, has_right); qb = pd_load_row_at(p + 5 * srp, has_left, has_right); }
        else { qa = pd_load_row(S, srp, 2 * y + 5, Hs, xs, Ws); qb = pd_load_row(S, srp, 2 * y + 6, Hs, xs, Ws); }
        if (y + 1 < y1) {
            pd_hsum_row(qc, r5); pd_hsum_row(qd, r6);
            vstore(drow + drp, r2, r3, r4, r5, r6);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) { r0[k] = r4[k]; r1[k] = r5[k]; r2[k] = r6[k]; }
        p += 4 * srp; drow += 2 * drp;
    }
}

// ---------------------------------------------------------------------------------------------
// a3, TMA form (sm_100a): the source rows of a CTA's segment arrive in shared memory by cp.async.bulk.tensor
// (one producer lane, NST-deep ring of 8-row stages, mbarrier full/empty handshake); the 4 consumer warps keep
// the same arithmetic as k_pyrdown_stream (a thread owns 8 output columns and slides down, DP4A horizontal
// sums packed two per register, five rows of them in registers), but never issue a global load: no
// long-scoreboard stalls, the loads of a segment run NST stages ahead of the arithmetic.
//   tensor map: 3D (W, H, B) uint8, box 256 x 8 x 1, out-of-bounds elements zero-filled.
//   CTA = (image b, column strip of 1920 source columns, PT_SEG output rows).  Stage j holds source rows
//   [2 y0 - 3 + 8 j, +8): one dummy row, the three prologue rows 2y0-2..2y0, then whole row pairs, so a
//   stage boundary never splits the two source rows of an output row.
//   REFLECT_101: left/right like the streaming kernel (resolved in registers by the edge threads); rows
//   above/below the image (first/last segment only) are read with ordinary mirrored loads.
// ---------------------------------------------------------------------------------------------
constexpr int PT_ROWS = 8, PT_NST = 3, PT_SEG_MAX = 46, PT_STRIP = 1920, PT_MAXW = 4;      // up to 4 consumer warps + 1 producer warp

__global__ void __launch_bounds__((PT_MAXW + 1) * 32, 6)
k_pyrdown_tma(const __grid_constant__ CUtensorMap map, const uint8_t *__restrict__ src, long long srp, long long sip,
              uint8_t *__restrict__ dst, long long drp, long long dip, int Hs, int Ws, int Hd, int Wd, int n_boxes,
              int strips, int segs_per_image, int n_segments, int PT_SEG)
{
    extern __shared__ __align__(1024) unsigned char pt_smem[];
    __shared__ __align__(8) unsigned long long bars[2 * PT_NST];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cw = (blockDim.x >> 5) - 1;                                  // consumer warps; the last warp is the producer
    const int PT_STAGE_BYTES = n_boxes * 256 * PT_ROWS;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[PT_NST]);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < PT_NST; i++) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, cw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t buf0 = smem_u32(pt_smem);
    // PERSISTENT: this CTA owns segments blockIdx.x, blockIdx.x + gridDim.x, ...; segment = (image, strip, PT_SEG output rows).
    // The producer runs up to PT_NST stages ahead ACROSS segment boundaries, so the ring never drains between segments.
    auto seg_geom = [&](int g, int &b, int &x0, int &y0, int &y1, int &n_stages) {
        const int per_img = strips * segs_per_image;
        b = g / per_img;
        const int r = g - b * per_img, sx = r % strips, sy = r / strips;
        x0 = sx * PT_STRIP;
        y0 = sy * PT_SEG; y1 = min(Hd, y0 + PT_SEG);
        n_stages = (2 * (y1 - y0) + 4 + PT_ROWS - 1) / PT_ROWS;           // dummy row + rows 2y0-2 .. 2(y1-1)+2
    };

    if (warp == cw) {
        if (lane == 0) {
            int js = 0;                                                    // stages issued by this CTA so far
            for (int g = blockIdx.x; g < n_segments; g += gridDim.x) {
                int b, x0, y0, y1, n_stages;
                seg_geom(g, b, x0, y0, y1, n_stages);
                for (int j = 0; j < n_stages; j++, js++) {
                    const int st = js % PT_NST;
                    if (js >= PT_NST) mbar_wait(empty0 + 8 * st, ((js / PT_NST) - 1) & 1);
                    mbar_expect_tx(full0 + 8 * st, (uint32_t)n_boxes * 256u * PT_ROWS);
                    for (int k = 0; k < n_boxes; k++)
                        tma_load_3d(buf0 + st * PT_STAGE_BYTES + k * 256 * PT_ROWS, &map, x0 - 16 + 256 * k, 2 * y0 - 3 + PT_ROWS * j, b,
                                    full0 + 8 * st);
                }
            }
        }
        return;
    }

    // ---- consumers: a thread owns 16 source / 8 output columns of the strip ----
    // a box is stored as PT_ROWS rows of 256 bytes; source column c of the strip (relative to x0 - 16) lives in box c >> 8
    const int c_main = 16 + 16 * tid;
    const int off_main = (c_main >> 8) * 256 * PT_ROWS + (c_main & 255);
    const int off_left = ((c_main - 4) >> 8) * 256 * PT_ROWS + ((c_main - 4) & 255);
    const int off_right = ((c_main + 16) >> 8) * 256 * PT_ROWS + ((c_main + 16) & 255);
    int js = 0;
    for (int g = blockIdx.x; g < n_segments; g += gridDim.x) {
        int b, x0, y0, y1, n_stages;
        seg_geom(g, b, x0, y0, y1, n_stages);
        const int base = 2 * y0 - 3;
        const int xs = x0 + 16 * tid;
        const bool active = xs < Ws && xs < x0 + PT_STRIP;
        const bool has_left = xs > 0, has_right = xs + 16 < Ws;
        const bool interior = base + 1 >= 0 && base + PT_ROWS * n_stages <= Hs;     // every row the segment reads is in the image
        const uint8_t *S = src + (long long)b * sip;
        uint8_t *drow = dst + (long long)b * dip + (long long)y0 * drp + (xs >> 1);
        auto row_fast = [&](const unsigned char *row) -> PdRaw {
            PdRaw q;
            q.v = *reinterpret_cast<const uint4 *>(row + off_main);
            q.l = has_left ? *reinterpret_cast<const uint32_t *>(row + off_left) : __byte_perm(q.v.x, 0u, 0x1200);
            q.r = has_right ? (uint32_t)row[off_right] : (q.v.w >> 16) & 255u;
            return q;
        };
        auto row_any = [&](const unsigned char *row, int t) -> PdRaw {     // mirrored rows: first / last segment of an image
            if (t < 0 || t >= Hs) return pd_load_row(S, srp, t, Hs, xs, Ws);
            return row_fast(row);
        };
        auto vstore = [&](const uint32_t (&a)[4], const uint32_t (&bb)[4], const uint32_t (&c)[4], const uint32_t (&d)[4],
                          const uint32_t (&e)[4]) {
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                o[k] = ((a[k] + e[k] + 0x00800080u + 4u * (bb[k] + d[k]) + 6u * c[k]) >> 8) & 0x00ff00ffu;
            *reinterpret_cast<uint2 *>(drow) = make_uint2(__byte_perm(o[0], o[1], 0x6420), __byte_perm(o[2], o[3], 0x6420));
            drow += drp;
        };
        uint32_t r0[4], r1[4], r2[4], r3[4], r4[4];
        int y = y0;
        for (int j = 0; j < n_stages; j++, js++) {
            const int st = js % PT_NST;
            mbar_wait(full0 + 8 * st, (js / PT_NST) & 1);
            const unsigned char *stage = pt_smem + st * PT_STAGE_BYTES;
            const int t0 = base + PT_ROWS * j;
            if (active) {
                if (interior && j > 0 && y + 4 <= y1) {                    // steady state: four row pairs, no tests
#pragma unroll
                    for (int rr = 0; rr < PT_ROWS; rr += 2) {
                        pd_hsum_row(row_fast(stage + rr * 256), r3);
                        pd_hsum_row(row_fast(stage + (rr + 1) * 256), r4);
                        vstore(r0, r1, r2, r3, r4);
#pragma unroll
                        for (int k = 0; k < 4; k++) { r0[k] = r2[k]; r1[k] = r3[k]; r2[k] = r4[k]; }
                    }
                    y += 4;
                } else {
                    int r = 0;
                    if (j == 0) {                                          // rows 1..3 of stage 0: the prologue rows 2y0-2 .. 2y0
                        pd_hsum_row(row_any(stage + 256, t0 + 1), r0);
                        pd_hsum_row(row_any(stage + 512, t0 + 2), r1);
                        pd_hsum_row(row_any(stage + 768, t0 + 3), r2);
                        r = 4;
                    }
#pragma unroll 1
                    for (int rr = r; rr < PT_ROWS && y < y1; rr += 2) {
                        pd_hsum_row(row_any(stage + rr * 256, t0 + rr), r3);
                        pd_hsum_row(row_any(stage + (rr + 1) * 256, t0 + rr + 1), r4);
                        vstore(r0, r1, r2, r3, r4);
#pragma unroll
                        for (int k = 0; k < 4; k++) { r0[k] = r2[k]; r1[k] = r3[k]; r2[k] = r4[k]; }
                        y++;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * st);
        }
    }
}

// 1 = launched, 0 = geometry not eligible (caller falls back to the streaming kernel), < 0 = error
static int launch_pyrdown_tma(const lkf_image *src, const lkf_image *dst, cudaStream_t s)
{
    if (src->width % 16 || src->width < 32 || src->height < 8 || !aligned_to(src, 16) || !aligned_to(dst, 8)) return 0;
    if (src->batch > 1 && src->image_pitch % 16) return 0;
    lkf_encode_tiled_fn enc = encode_tiled();
    if (!enc) return 0;
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)src->width, (cuuint64_t)src->height, (cuuint64_t)src->batch};
    const cuuint64_t gstr[2] = {(cuuint64_t)src->row_pitch, (cuuint64_t)(src->batch > 1 ? src->image_pitch : src->row_pitch * src->height)};
    const cuuint32_t box[3] = {256, PT_ROWS, 1}, estr[3] = {1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, src->data, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    const int strips = (src->width + PT_STRIP - 1) / PT_STRIP;
    const int strip_w = src->width < PT_STRIP ? src->width : PT_STRIP;
    const int n_boxes = (strip_w + 32 + 255) / 256;                        // columns x0-16 .. x0+strip+16
    const int cw = (strip_w / 16 + 31) / 32;                               // consumer warps: 16 source columns per thread
    const int nt = (cw + 1) * 32;
    const size_t smem = (size_t)PT_NST * n_boxes * 256 * PT_ROWS + 1024;
    LKF_FUNC_ATTRS(LKF_CUDA(cudaFuncSetAttribute(k_pyrdown_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pyrdown_tma, nt, smem);
    if (per_sm < 1) return 0;
    const long long resident = (long long)per_sm * sms;                    // persistent: every CTA is resident
    // output rows per segment: 2 + 4m (stage 0 yields two rows, every later stage four); as long as possible while the
    // segments still fill the resident CTAs about three times over
    int seg = PT_SEG_MAX;
    while (seg > 6 && (long long)strips * ((dst->height + seg - 1) / seg) * dst->batch < 3 * resident) seg -= 4;
    const int segs = (dst->height + seg - 1) / seg;
    const long long n_seg = (long long)strips * segs * dst->batch;
    const int grid = (int)(n_seg < resident ? n_seg : resident);
    k_pyrdown_tma<<<grid, nt, smem, s>>>(map, (const uint8_t *)src->data, src->row_pitch, src->image_pitch, (uint8_t *)dst->data,
                                         dst->row_pitch, dst->image_pitch, src->height, src->width, dst->height, dst->width, n_boxes,
                                         strips, segs, (int)n_seg, seg);
    return 1;
}

bool same_geometry(const lkf_image *a, const lkf_image *b)
{
    return a->batch == b->batch && a->height == b->height && a->width == b->width;
}

}  // namespace

LKF_EXPORT int lkf_bgr2gray_u8(const lkf_image *bgr, const lkf_image *gray, lkf_stream stream)
{
    if (!img_ok(bgr, 3, 1) || !img_ok(gray, 1, 1) || !same_geometry(bgr, gray)) return LKF_ERR_ARG;
    Img s = to_img(bgr), d = to_img(gray);
    int al16 = aligned_to(bgr, 16) && aligned_to(gray, 16), al4 = aligned_to(bgr, 4) && aligned_to(gray, 4);
    dim3 grid((s.W + 16 * NT - 1) / (16 * NT), s.H, s.B);
    k_bgr2gray<<<grid, NT, 0, (cudaStream_t)stream>>>(s, d, al16, al4);
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}

LKF_EXPORT int lkf_gaussian_blur3_u8(const lkf_image *src, const lkf_image *dst, lkf_stream stream)
{
    if (!img_ok(src, 1, 1) || !img_ok(dst, 1, 1) || !same_geometry(src, dst) || src->data == dst->data) return LKF_ERR_ARG;
    if (stream_ok(src, dst)) {
        launch_blur3_stream<RowU8>(src, dst, (cudaStream_t)stream);
        LKF_LAUNCH_CHECK(nullptr);
        return LKF_OK;
    }
    Img d = to_img(dst);
    SrcU8 s{(const uint8_t *)src->data, src->row_pitch, src->height, src->width, aligned_to(src, 4)};
    dim3 grid((d.W + BT_W - 1) / BT_W, (d.H + BT_H - 1) / BT_H, d.B);
    k_blur3<SrcU8><<<grid, NT, 0, (cudaStream_t)stream>>>(s, src->image_pitch, d, aligned_to(dst, 4));
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}

LKF_EXPORT int lkf_process_img_u8(const lkf_image *bgr, const lkf_image *dst, lkf_stream stream)
{
    if (!img_ok(bgr, 3, 1) || !img_ok(dst, 1, 1) || !same_geometry(bgr, dst)) return LKF_ERR_ARG;
    if (stream_ok(bgr, dst)) {
        launch_blur3_stream<RowBGR>(bgr, dst, (cudaStream_t)stream);
        LKF_LAUNCH_CHECK(nullptr);
        return LKF_OK;
    }
    Img d = to_img(dst);
    SrcBGR s{(const uint8_t *)bgr->data, bgr->row_pitch, bgr->height, bgr->width, aligned_to(bgr, 4)};
    dim3 grid((d.W + BT_W - 1) / BT_W, (d.H + BT_H - 1) / BT_H, d.B);
    k_blur3<SrcBGR><<<grid, NT, 0, (cudaStream_t)stream>>>(s, bgr->image_pitch, d, aligned_to(dst, 4));
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}

LKF_EXPORT int lkf_pyr_down_u8(const lkf_image *src, const lkf_image *dst, lkf_stream stream)
{
    if (!img_ok(src, 1, 1) || !img_ok(dst, 1, 1) || src->batch != dst->batch ||
        dst->height != (src->height + 1) / 2 || dst->width != (src->width + 1) / 2)
        return LKF_ERR_ARG;
    // TMA form for wide levels (measured, 65 frames of 1920x1080 -> 960x540: 34.8-35.2 us vs 39.8 us for the streaming kernel;
    // 960x540 -> 480x270: 14.0 vs 13.5 us, so narrow levels keep the streaming kernel).  LKF_PYR_TMA=0/1 forces one form
    // (developer switch for A/B timing, read once).
    static const int force_tma = getenv("LKF_PYR_TMA") ? atoi(getenv("LKF_PYR_TMA")) : -1;
    if (force_tma == 1 || (force_tma < 0 && src->width >= 1280)) {
        const int rc = launch_pyrdown_tma(src, dst, (cudaStream_t)stream);
        if (rc < 0) return rc;
        if (rc == 1) { LKF_LAUNCH_CHECK(nullptr); return LKF_OK; }
    }
    if (src->width % 16 == 0 && src->width >= 32 && src->height >= 4 && aligned_to(src, 16) && aligned_to(dst, 8)) {
        // rows per CTA: 16, fewer when that leaves the GPU short of about one wave of resident warps (small levels)
        const int nt = stream_threads(dst->width / 8);
        const int gx = (dst->width / 8 + nt - 1) / nt;
        int R = 16;
        while (R > 4 && (long long)gx * ((dst->height + R - 1) / R) * dst->batch * (nt / 32) < 148LL * 28) R >>= 1;
        dim3 g(gx, (dst->height + R - 1) / R, dst->batch);
        k_pyrdown_stream<<<g, nt, 0, (cudaStream_t)stream>>>((const uint8_t *)src->data, src->row_pitch, src->image_pitch,
                                                                (uint8_t *)dst->data, dst->row_pitch, dst->image_pitch, src->height,
                                                                src->width, dst->height, dst->width, R);
        LKF_LAUNCH_CHECK(nullptr);
        return LKF_OK;
    }
    Img d = to_img(dst);
    SrcU8 s{(const uint8_t *)src->data, src->row_pitch, src->height, src->width, aligned_to(src, 4)};
    dim3 grid((d.W + PD_W - 1) / PD_W, (d.H + PD_H - 1) / PD_H, d.B);
    k_pyrdown<<<grid, NT, 0, (cudaStream_t)stream>>>(s, src->image_pitch, d, aligned_to(dst, 4));
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}

LKF_EXPORT int lkf_scharr_s16(const lkf_image *src, const lkf_image *dst, lkf_stream stream)
{
    if (!img_ok(src, 1, 1) || !img_ok(dst, 2, 2) || !same_geometry(src, dst) || !aligned_to(dst, 4)) return LKF_ERR_ARG;
    Img d = to_img(dst);
    SrcU8 s{(const uint8_t *)src->data, src->row_pitch, src->height, src->width, aligned_to(src, 4)};
    dim3 grid((d.W + BT_W - 1) / BT_W, (d.H + BT_H - 1) / BT_H, d.B);
    k_scharr<<<grid, NT, 0, (cudaStream_t)stream>>>(s, src->image_pitch, d, aligned_to(dst, 16));
    LKF_LAUNCH_CHECK(nullptr);
    return LKF_OK;
}

LKF_EXPORT int lkf_effective_max_level(int height, int width, int win_w, int win_h, int max_level)
{
    int w = width, h = height;
    for (int level = 0; level <= max_level; level++) {
        w = (w + 1) / 2; h = (h + 1) / 2;
        if (w <= win_w || h <= win_h) return level;
    }
    return max_level;
}

/* a1+a2+a3 for a whole pyramid: levels[0] = process_img(bgr), levels[l] = pyrDown(levels[l-1]) -- one ABI call, the
 * streaming kernels back to back.  (A single fused pass -- warp-sliding, all three levels in registers -- was built and
 * measured in round 1: bit-exact but SLOWER, 200 us vs 108 + 76 us per 64 frames, because process_img is already
 * issue-bound on B200 (18 instructions per pixel at 4.9 TB/s) and the halo rows/lanes of the fused form add more
 * instructions than the saved 1.25 P of traffic is worth.) */
LKF_EXPORT int lkf_process_pyramid_u8(const lkf_image *bgr, const lkf_image *levels, int n_levels, lkf_stream stream)
{
    if (!bgr || !levels || n_levels < 1 || n_levels > LKF_MAX_LEVELS) return LKF_ERR_ARG;
    if (!img_ok(bgr, 3, 1) || !img_ok(&levels[0], 1, 1) || !same_geometry(bgr, &levels[0])) return LKF_ERR_ARG;
    for (int l = 1; l < n_levels; l++)
        if (!img_ok(&levels[l], 1, 1) || levels[l].batch != bgr->batch || levels[l].height != (levels[l - 1].height + 1) / 2 ||
            levels[l].width != (levels[l - 1].width + 1) / 2)
            return LKF_ERR_ARG;
    int rc = lkf_process_img_u8(bgr, &levels[0], stream);
    for (int l = 1; l < n_levels && rc == LKF_OK; l++) rc = lkf_pyr_down_u8(&levels[l - 1], &levels[l], stream);
    return rc;
}
