// probe: which nvJPEG backends exist on this GPU and how fast they decode a batch of 1080p baseline JPEGs to BGR (interleaved)
#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb"); if (!f) { printf("no file\n"); return 1; }
    fseek(f, 0, SEEK_END); size_t n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> jpg(n); fread(jpg.data(), 1, n, f); fclose(f);
    const int B = argc > 2 ? atoi(argv[2]) : 64, W = 1920, H = 1080;
    const char *names[] = {"DEFAULT", "HYBRID", "GPU_HYBRID", "HARDWARE"};
    for (int be = 0; be <= 3; be++) {
        nvjpegHandle_t h; nvjpegJpegState_t st;
        nvjpegStatus_t rc = nvjpegCreateEx((nvjpegBackend_t)be, NULL, NULL, 0, &h);
        if (rc != NVJPEG_STATUS_SUCCESS) { printf("%-10s create failed: status %d\n", names[be], (int)rc); continue; }
        nvjpegJpegStateCreate(h, &st);
        rc = nvjpegDecodeBatchedInitialize(h, st, B, 1, NVJPEG_OUTPUT_BGRI);
        if (rc != NVJPEG_STATUS_SUCCESS) { printf("%-10s batched init failed: status %d\n", names[be], (int)rc); nvjpegJpegStateDestroy(st); nvjpegDestroy(h); continue; }
        std::vector<nvjpegImage_t> out(B);
        std::vector<const unsigned char *> data(B, jpg.data());
        std::vector<size_t> len(B, n);
        unsigned char *buf; cudaMalloc(&buf, (size_t)B * W * H * 3);
        for (int i = 0; i < B; i++) { for (int c = 0; c < 4; c++) { out[i].channel[c] = NULL; out[i].pitch[c] = 0; } out[i].channel[0] = buf + (size_t)i * W * H * 3; out[i].pitch[0] = W * 3; }
        cudaStream_t s; cudaStreamCreate(&s);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        rc = nvjpegDecodeBatched(h, st, data.data(), len.data(), out.data(), s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (rc != NVJPEG_STATUS_SUCCESS || e != cudaSuccess) { printf("%-10s decode failed: status %d cuda %s\n", names[be], (int)rc, cudaGetErrorString(e)); }
        else {
            cudaEventRecord(a, s);
            const int reps = 5;
            for (int r = 0; r < reps; r++) nvjpegDecodeBatched(h, st, data.data(), len.data(), out.data(), s);
            cudaEventRecord(b, s); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            unsigned char px[6]; cudaMemcpy(px, buf + ((size_t)540 * W + 960) * 3, 6, cudaMemcpyDeviceToHost);
            printf("%-10s ok: %d x %zu-byte 1080p JPEGs in %.2f ms per batch = %.0f images/s  (pixel %d %d %d)\n", names[be], B, n, ms / reps, B * reps / (ms * 1e-3), px[0], px[1], px[2]);
        }
        cudaFree(buf); nvjpegJpegStateDestroy(st); nvjpegDestroy(h);
    }
    return 0;
}
