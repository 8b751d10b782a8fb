// standalone probe: one warp loads a (bw x bh x 1) box by TMA from a (W,H,B) u8 tensor; descriptor = maps.m[idx] of a 1 KB param
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include "../lk-optical-flow-method_b200/csrc/tma.cuh"
struct Maps { CUtensorMap m[8]; };
__global__ void __launch_bounds__(32, 28) k(int pad0, int pad1, double pad2, const __grid_constant__ Maps maps, int idx, int x, int y, int b, int bytes, int dst_off, int bar_off, unsigned *out, int mode)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x;
    const uint32_t bar = smem_u32(sm + bar_off);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    if (lane == 0) { if (mode == 1) fence_proxy_async(); if (mode == 2) asm volatile("fence.proxy.async;" ::: "memory"); mbar_expect_tx(bar, bytes); tma_load_3d(smem_u32(sm + dst_off), &maps.m[idx], x, y, b, bar); }
    mbar_wait(bar, 0);
    __syncwarp();
    unsigned s = 0;
    for (int i = lane; i < bytes; i += 32) s += sm[dst_off + i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[blockIdx.x] = s;
}
__global__ void __launch_bounds__(64) k2(const __grid_constant__ Maps maps, int idx, int x, int y, int b, int bytes, int dst_off, int bar_off, unsigned *out, int mode)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long sbar[1];
    const int lane = threadIdx.x;
    if (lane >= 32) return;
    const uint32_t bar = mode & 4 ? smem_u32(&sbar[0]) : smem_u32(sm + bar_off);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    if (lane == 0) { mbar_expect_tx(bar, bytes); if (mode & 8) tma_load_3d(smem_u32(sm + dst_off), &maps.m[0], x, y, b, bar); else tma_load_3d(smem_u32(sm + dst_off), &maps.m[idx], x, y, b, bar); }
    mbar_wait(bar, 0);
    __syncwarp();
    unsigned s = 0;
    for (int i = lane; i < bytes; i += 32) s += sm[dst_off + i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[blockIdx.x] = s;
}
int main(int argc, char **argv)
{
    int bw = atoi(argv[1]), bh = atoi(argv[2]), idx = atoi(argv[3]), B = atoi(argv[4]), dst_off = atoi(argv[5]), bar_off = atoi(argv[6]);
    int W = 1920, H = 1080;
    uint8_t *d; cudaMalloc(&d, (size_t)W * H * B);
    uint8_t *h = (uint8_t *)malloc((size_t)W * H * B);
    for (size_t i = 0; i < (size_t)W * H * B; i++) h[i] = (uint8_t)(i * 7 + (i >> 11));
    cudaMemcpy(d, h, (size_t)W * H * B, cudaMemcpyHostToDevice);
    Maps maps; memset(&maps, 0, sizeof(maps));
    lkf_encode_tiled_fn enc = encode_tiled();
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B}, gstr[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&maps.m[idx], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     (atoi(argv[7]) & 16) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d  ", (int)r);
    unsigned *out; cudaMalloc(&out, 4096);
    int x = argc > 8 ? atoi(argv[8]) : 100, y = 50, b = B - 1;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    int mode = atoi(argv[7]);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    if (mode >= 4) k2<<<4, (mode & 32) ? 64 : 32, 16384>>>(maps, idx, x, y, b, bw * bh, dst_off, bar_off, out, mode);
    else k<<<4, 32, 16384>>>(0, 0, 0.0, maps, idx, x, y, b, bw * bh, dst_off, bar_off, out, mode);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned got = 0; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
    unsigned want = 0;
    for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++) want += h[(size_t)b * W * H + (size_t)(y + j) * W + x + i];
    printf("bw=%d bh=%d idx=%d B=%d dst=%d bar=%d: %s got=%u want=%u\n", bw, bh, idx, B, dst_off, bar_off, cudaGetErrorString(e), got, want);
    return 0;
}
