// v1_texture.cu -- the two texture-side pieces the reference runs outside its KTX2 path:
//
//  * the V1 player's frame counter: every video frame carries its frame number as 16 black / white cells in the bottom-left corner
//    (painted by example/texture_encoder.py:59-63, read back by drawVideoAndGetCurrentFrameNumber, src/V1/player.ts:305-334: the
//    128 x 4 pixel strip is scaled to 16 x 1, bit i = round(R_i / 255), frame = max(sum(bit_i << i) - 1, 0)).  Here one warp reads the
//    strip of one decoded RGBA frame (lane i averages cell i: the canvas down-scale is a box filter over whole cells) -- for frames
//    already in HBM (NVDEC output) or in host memory.  The H.264 decode itself is not part of this library.
//  * the 'etc2' texture target of the V2 player (src/V2/player.ts:338-356,454-470): one raw RGB-ETC2 block file per frame that the
//    reference hands to the GPU as it is (CompressedTexture, RGB_ETC2_Format).  The equivalent here is a validated batched upload:
//    every file must hold exactly ceil(w/4) * ceil(h/4) * 8 bytes.
#include <string.h>
#include "uvol_ctx.h"

namespace {
__global__ void __launch_bounds__(128) k_frame_counter(const uint8_t *rgba, int n, int width, int height, int cell, int cells, int32_t *out) {
    const int fi = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (fi >= n) return;
    const uint8_t *img = rgba + (size_t)fi * width * height * 4;
    const int rows = cell / 2;                                  // encoderWindowHeight = encoderWindowSize / 2 (src/V1/player.ts:307)
    uint32_t sum = 0;
    if (lane < cells) for (int y = 0; y < rows; y++) for (int x = 0; x < cell; x++) sum += img[(((size_t)(height - rows + y)) * width + lane * cell + x) * 4];
    const uint32_t bit = lane < cells && 2u * sum >= 255u * (uint32_t)(rows * cell) ? 1u : 0u;          // Math.round(mean / 255)
    const unsigned mask = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) { const int v = (int)(mask & ((cells >= 32 ? 0u : (1u << cells)) - 1u)) - 1; out[fi] = v < 0 ? 0 : v; }
}
}  // namespace

// frames: n RGBA8 images of width x height, back to back, in device memory (frames_on_device != 0) or host memory.  window_size /
// byte_length: the player's encoderWindowSize (8) and encoderByteLength (16) (src/Player.ts:47-48).  out[i] = frame number of image i.
extern "C" int uvol_v1_frame_numbers(uvol_ctx *ctx, const uint8_t *frames, int frames_on_device, int n, int width, int height, int window_size, int byte_length, int32_t *out) {
    if (!ctx || !frames || !out || n < 0 || width <= 0 || height <= 0 || window_size < 2 || byte_length < 1 || byte_length > 31) return UVOL_ERR_ARG;
    if ((long long)window_size * byte_length > width || window_size / 2 > height) return UVOL_ERR_ARG;
    if (n == 0) return UVOL_OK;
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n * width * height * 4;
    const uint8_t *d = frames;
    if (!frames_on_device) {          // only the bottom rows of each image are needed: copy those (2D copy, one row block per image)
        const size_t strip = (size_t)(window_size / 2) * width * 4;
        UVOL_CUDA(ctx, ctx->d_flush.reserve(strip * (size_t)n + 256));
        UVOL_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_flush.p, strip, frames + (size_t)(height - window_size / 2) * width * 4, (size_t)width * height * 4, strip, (size_t)n, cudaMemcpyHostToDevice, ctx->s0));
        d = (const uint8_t *)ctx->d_flush.p; height = window_size / 2;
    }
    (void)bytes;
    UVOL_CUDA(ctx, ctx->d_ccounts.reserve(4 * (size_t)n + 256)); UVOL_CUDA(ctx, ctx->h_ccounts.reserve(4 * (size_t)n + 256));
    k_frame_counter<<<(unsigned)((n + 3) / 4), 128, 0, ctx->s0>>>(d, n, width, height, window_size, byte_length, (int32_t *)ctx->d_ccounts.p);
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_ccounts.p, ctx->d_ccounts.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->s0));
    UVOL_CUDA(ctx, cudaStreamSynchronize(ctx->s0));
    UVOL_CUDA(ctx, cudaGetLastError());
    memcpy(out, ctx->h_ccounts.p, 4 * (size_t)n);
    return UVOL_OK;
}

// n raw block-compressed texture files of one resolution (the 'etc2' target: RGB ETC2, 8 bytes per 4x4 block).  Each file is checked
// against the block count of width x height; good ones are uploaded back to back (UVOL_MEM_DEVICE) or copied into pinned host memory
// (UVOL_MEM_HOST); out[i].data points at file i's blocks, out[i].format = UVOL_TEX_ETC2_RGB.
extern "C" int uvol_upload_etc2_batch(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int width, int height, int memory, uvol_texture *out) {
    if (!ctx || !out || n < 0 || (n > 0 && (!data || !size)) || width <= 0 || height <= 0 || width > 16384 || height > 16384) return UVOL_ERR_ARG;
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t per = (uint64_t)((width + 3) / 4) * ((height + 3) / 4) * 8;
    uint64_t total = 0;
    for (int i = 0; i < n; i++) { memset(&out[i], 0, sizeof out[i]); out[i].status = (data[i] && size[i] == per) ? UVOL_OK : (data[i] && size[i] < per ? UVOL_ERR_TRUNCATED : UVOL_ERR_CORRUPT); if (!out[i].status) total += align_up(per, 128); }
    if (!total) return UVOL_OK;
    UVOL_CUDA(ctx, ctx->ph_tout->reserve(total + 256));
    uint8_t *h = (uint8_t *)ctx->ph_tout->p; uint64_t off = 0;
    for (int i = 0; i < n; i++) if (!out[i].status) { memcpy(h + off, data[i], per); off += align_up(per, 128); }
    uint8_t *base = h;
    if (memory == UVOL_MEM_DEVICE) {
        UVOL_CUDA(ctx, ctx->d_out_tex.reserve(total + 256));
        UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_out_tex.p, h, total, cudaMemcpyHostToDevice, ctx->s2));
        UVOL_CUDA(ctx, cudaStreamSynchronize(ctx->s2));
        base = (uint8_t *)ctx->d_out_tex.p;
    }
    off = 0;
    for (int i = 0; i < n; i++) if (!out[i].status) {
        out[i].width = (uint32_t)width; out[i].height = (uint32_t)height; out[i].layers = 1; out[i].levels = 1; out[i].format = UVOL_TEX_ETC2_RGB; out[i].data = base + off; out[i].bytes = per;
        off += align_up(per, 128);
    }
    return UVOL_OK;
}
