// abi.cu -- library-level entry points: version, error strings, workspace handle.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <atomic>

static thread_local char g_last_error[512] = "";
static std::atomic<long long> g_launches{0};

void lkf_set_error(const char *what, cudaError_t e)
{
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();   // clear the sticky-free error so later calls report their own failures
}

void lkf_count_launch(lkf_ctx *ctx)
{
    if (ctx) ctx->launches++;
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

// Scratch of a ctx: grow-only.  A ctx serves ONE stream at a time (calls on one ctx are not re-entrant and its scratch is not
// ordered across streams).  A pointer handed out while `stream` is being captured is baked into the CUDA graph, so that slot
// is pinned from then on: it is never freed, and a later request that would need it to grow is refused
// (LKF_ERR_UNSUPPORTED) instead of leaving the graph with a dangling pointer -- size everything with one eager call before
// capturing, and give every captured pipeline its own ctx.  Growth itself (cudaFree + cudaMalloc) is refused during capture.
int lkf_scratch(lkf_ctx *ctx, int slot, size_t bytes, void **out, cudaStream_t stream)
{
    if (!ctx || slot < 0 || slot >= 8) return LKF_ERR_ARG;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (stream) LKF_CUDA(cudaStreamIsCapturing(stream, &cs));
    const bool capturing = cs != cudaStreamCaptureStatusNone;
    if (ctx->scratch_bytes[slot] < bytes) {
        if (capturing || ctx->scratch_pinned[slot]) {
            snprintf(g_last_error, sizeof(g_last_error),
                     "workspace scratch %d would have to grow from %zu to %zu bytes %s: run the call once eagerly at full size "
                     "before capturing, and use one lkf_ctx per captured pipeline", slot, ctx->scratch_bytes[slot], bytes,
                     capturing ? "during stream capture" : "but a captured CUDA graph holds its address");
            return LKF_ERR_UNSUPPORTED;
        }
        // growth synchronises (cudaFree), steady state does not
        if (ctx->scratch[slot]) LKF_CUDA(cudaFree(ctx->scratch[slot]));
        ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
        if (e != cudaSuccess) { lkf_set_error("cudaMalloc(scratch)", e); return LKF_ERR_NOMEM; }
        ctx->scratch_bytes[slot] = want;
    }
    if (capturing) ctx->scratch_pinned[slot] = true;
    *out = ctx->scratch[slot];
    return LKF_OK;
}

LKF_EXPORT int lkf_abi_version(void) { return LKF_ABI_VERSION; }

LKF_EXPORT int lkf_struct_size(int which)
{
    switch (which) {
        case 0: return (int)sizeof(lkf_image);
        case 1: return (int)sizeof(lkf_lk_params);
        case 2: return (int)sizeof(lkf_gftt_params);
        case 3: return (int)sizeof(lkf_vp_params);
        case 4: return (int)sizeof(lkf_vp_state);
        default: return -1;
    }
}

LKF_EXPORT const char *lkf_strerror(int status)
{
    switch (status) {
        case LKF_OK: return "ok";
        case LKF_ERR_ARG: return "bad argument";
        case LKF_ERR_CUDA: return "CUDA runtime error";
        case LKF_ERR_UNSUPPORTED: return "unsupported configuration";
        case LKF_ERR_NOMEM: return "out of device memory";
        case LKF_ERR_OVERFLOW: return "output capacity too small";
        default: return "unknown status";
    }
}

LKF_EXPORT const char *lkf_last_error(void) { return g_last_error; }

LKF_EXPORT int lkf_ctx_create(int device, lkf_ctx **out)
{
    if (!out) return LKF_ERR_ARG;
    int ndev = 0;
    LKF_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return LKF_ERR_ARG;
    cudaDeviceProp prop;
    LKF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        snprintf(g_last_error, sizeof(g_last_error), "device %d is sm_%d%d; this library is built for sm_100a only",
                 device, prop.major, prop.minor);
        return LKF_ERR_UNSUPPORTED;
    }
    lkf_ctx *c = (lkf_ctx *)calloc(1, sizeof(lkf_ctx));
    if (!c) return LKF_ERR_NOMEM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return LKF_OK;
}

LKF_EXPORT void lkf_ctx_destroy(lkf_ctx *ctx)
{
    if (!ctx) return;
    for (int i = 0; i < 8; i++)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    free(ctx);
}

LKF_EXPORT int64_t lkf_ctx_launch_count(const lkf_ctx *ctx)
{
    return ctx ? ctx->launches : (int64_t)g_launches.load();
}
