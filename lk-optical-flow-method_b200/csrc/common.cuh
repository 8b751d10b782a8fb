// common.cuh -- shared helpers of liblkflow_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/lkflow_b200.h"

#define LKF_EXPORT extern "C" __attribute__((visibility("default")))

struct lkf_ctx {
    int device;
    int sm_count;
    int64_t launches;
    // grow-only scratch (GFTT candidates, sort keys, cell grid, per-image scalars)
    void *scratch[8];
    size_t scratch_bytes[8];
    bool scratch_pinned[8];     // handed out while a stream was capturing: a CUDA graph holds the raw pointer, never freed again
};

void lkf_set_error(const char *what, cudaError_t e);
int lkf_scratch(lkf_ctx *ctx, int slot, size_t bytes, void **out, cudaStream_t stream);

#define LKF_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) { lkf_set_error(#call, e__); return LKF_ERR_CUDA; } \
    } while (0)

#define LKF_LAUNCH_CHECK(ctx_)                                           \
    do {                                                                 \
        cudaError_t e__ = cudaPeekAtLastError();                         \
        if (e__ != cudaSuccess) { lkf_set_error("kernel launch", e__); return LKF_ERR_CUDA; } \
        lkf_count_launch(ctx_);                                          \
    } while (0)

void lkf_count_launch(lkf_ctx *ctx);   // ctx may be NULL -> global counter

// Function attributes (dynamic shared memory size, carveout) are per DEVICE: one flag word per kernel instantiation,
// bit d = "set on device d".  LKF_FUNC_ATTRS(flags, stmts) runs stmts once per device (benign race: the attribute calls
// are idempotent, the flag is published after them); cudaFuncSetAttribute is far too slow for the per-launch path.
#include <atomic>
#define LKF_FUNC_ATTRS(stmts)                                                         \
    do {                                                                              \
        static std::atomic<unsigned long long> done__{0};                             \
        int dev__ = 0;                                                                \
        LKF_CUDA(cudaGetDevice(&dev__));                                              \
        const unsigned long long bit__ = 1ull << (dev__ & 63);                        \
        if (!(done__.load(std::memory_order_acquire) & bit__)) {                      \
            stmts;                                                                    \
            done__.fetch_or(bit__, std::memory_order_release);                        \
        }                                                                             \
    } while (0)

// Device-side view of an lkf_image
struct Img {
    uint8_t *p;
    int B, H, W, C;
    long long rp, ip;
};
static inline Img to_img(const lkf_image *im)
{
    Img r;
    r.p = (uint8_t *)im->data; r.B = im->batch; r.H = im->height; r.W = im->width; r.C = im->channels;
    r.rp = im->row_pitch; r.ip = im->image_pitch;
    return r;
}
static inline bool img_ok(const lkf_image *im, int channels, int elem)
{
    return im && im->data && im->batch > 0 && im->height > 0 && im->width > 0 && im->channels == channels &&
           im->row_pitch >= (int64_t)im->width * channels * elem &&
           (im->batch == 1 || im->image_pitch >= im->row_pitch * im->height);
}
static inline bool aligned_to(const lkf_image *im, int a)
{
    return ((uintptr_t)im->data % a) == 0 && (im->row_pitch % a) == 0 && (im->batch == 1 || (im->image_pitch % a) == 0);
}

// gfedcb|abcdefgh|gfedcba, any i (multiple reflections), n >= 1
__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

__device__ __forceinline__ unsigned gray_of(unsigned b, unsigned g, unsigned r)
{
    // Y = (3735 B + 19235 G + 9798 R + 2^14) >> 15   (bit-exact vs cv2 4.13)
    return (3735u * b + 19235u * g + 9798u * r + (1u << 14)) >> 15;
}
