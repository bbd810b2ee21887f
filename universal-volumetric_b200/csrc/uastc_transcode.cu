// uastc_transcode.cu -- UASTC LDR 4x4 -> RGBA32 block kernel for sm_100a.
//
// Replaces ktx2File.transcodeImage(dst, mip, layer, face, RGBA32, ...) for KTX2 files whose DFD colour model is UASTC
// (src/lib/KTX2Loader.js:487,551-552).  UASTC carries no entropy coding: every 16-byte block decodes on its own, so the
// pass is element-parallel and HBM-bound (16 B in, 64 B out per block = 80 B of compulsory traffic).
// One thread per block; a warp covers 32 adjacent blocks of a block row, so every pixel-row store of the warp is 512
// contiguous bytes.  The block is held in four registers and consumed front to back with funnel shifts; all
// per-mode parameters, the partition patterns and the endpoint / weight unquantisation tables sit in shared memory
// (lanes of a warp look up different modes, which a __constant__ bank would serialise).  Control flow is table-driven
// rather than a switch per mode, so blocks of different modes in one warp mostly share instructions.
// Block format: see oracle/uastc_oracle.c (the CPU restatement this kernel is checked against bit for bit).
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include "uvol_ctx.h"
#include "uastc_core.h"
#include "bc7_core.h"
#include "astc_core.h"
#include "tma_bulk.h"

namespace {

__device__ __align__(16) uint32_t g_tables[sizeof(UastcShared) / 4];      // image of UastcShared, filled once per device by the launcher

// grid = (ceil(max blocks / (256 * UASTC_CHUNKS)), UASTC layer list): a CTA loads the tables once and decodes UASTC_CHUNKS x 256 blocks
#define UASTC_CHUNKS 8
// TMA = true: the 2.8 KB table image is staged by ONE bulk asynchronous copy (cp.async.bulk, completion on an mbarrier) issued by
// thread 0 while the other threads already fetch the per-layer constants; TMA = false: the word-by-word copy by all threads (kept
// for the A/B measurement, UVOL_NO_TMA=1).
template <bool TMA>
__global__ void __launch_bounds__(256) k_uastc_blocks(const Ktx2File *files, int32_t *status2, const uint8_t *blob, uint8_t *O, const uint32_t *layer_list) {
    __shared__ __align__(16) UastcShared T; __shared__ __align__(8) uint64_t bar;
    if (TMA) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_expect_tx(&bar, (uint32_t)sizeof(UastcShared)); bulk_g2s(&T, g_tables, (uint32_t)sizeof(UastcShared), &bar); }
    } else {
        for (uint32_t i = threadIdx.x; i < sizeof(UastcShared) / 4; i += 256) ((uint32_t *)&T)[i] = g_tables[i];
    }
    __shared__ struct { const uint8_t *src0; uint8_t *dst; uint32_t nblk, W, H, bxn, fi, skip; } K;      // per-layer constants, fetched once per CTA
    if (threadIdx.x == 0) {
        const uint32_t ll = layer_list[blockIdx.y], fi = ll >> 12, L = ll & 4095;
        const Ktx2File &f = files[fi];
        K.skip = f.status || !f.is_uastc; K.fi = fi;
        K.nblk = f.bx * f.by; K.W = f.width; K.H = f.height; K.bxn = f.bx;
        K.src0 = blob + f.file_off + f.level_off + (size_t)L * K.nblk * 16;
        K.dst = O + f.o_rgba + (size_t)L * f.width * f.height * 4;
    }
    __syncthreads();
    if (TMA) mbar_wait(&bar, 0);
    if (K.skip) return;
    const uint32_t nblk = K.nblk, W = K.W, H = K.H, bxn = K.bxn, fi = K.fi;
    const uint8_t *src0 = K.src0; uint8_t *dst = K.dst;
    const bool aligned = (((uintptr_t)src0) & 15) == 0, whole = (W & 3) == 0;
#pragma unroll 1
    for (uint32_t k = 0; k < UASTC_CHUNKS; k++) {
        const uint32_t bi = (blockIdx.x * UASTC_CHUNKS + k) * 256 + threadIdx.x;
        if (bi >= nblk) return;
        const uint8_t *src = src0 + (size_t)bi * 16;
        uint4 blk;
        if (aligned) blk = __ldcs((const uint4 *)src);
        else { uint32_t w[4]; for (int q = 0; q < 4; q++) w[q] = src[4 * q] | (src[4 * q + 1] << 8) | (src[4 * q + 2] << 16) | ((uint32_t)src[4 * q + 3] << 24); blk = make_uint4(w[0], w[1], w[2], w[3]); }
        uint32_t rows[4][4];
        if (!uastc_block(T, blk.x, blk.y, blk.z, blk.w, rows)) { status2[2 * fi] = UVOL_ERR_CORRUPT; continue; }      // like a failed transcodeImage: the whole segment is an error
        const uint32_t xb = bi % bxn, yb = bi / bxn;
        if (whole && xb * 4 + 4 <= W) {
#pragma unroll
            for (uint32_t y = 0; y < 4; y++) if (yb * 4 + y < H)
                __stcs((uint4 *)(dst + ((size_t)(yb * 4 + y) * W + xb * 4) * 4), make_uint4(rows[y][0], rows[y][1], rows[y][2], rows[y][3]));
        } else {
            for (uint32_t y = 0; y < 4 && yb * 4 + y < H; y++) for (uint32_t x = 0; x < 4 && xb * 4 + x < W; x++)
                *(uint32_t *)(dst + ((size_t)(yb * 4 + y) * W + xb * 4 + x) * 4) = rows[y][x];
        }
    }
}

// UASTC -> BC7 (UVOL_TEX_BC7) / UASTC -> ASTC 4x4 (UVOL_TEX_ASTC_4x4): same traversal of the blocks, 16 bytes out per block in block
// raster order (a warp stores 512 contiguous bytes); per-block logic in bc7_core.h / astc_core.h.  X = the target's table image.
__device__ __align__(16) uint32_t g_bc7_tables[sizeof(Bc7Shared) / 4];
__device__ __align__(16) uint32_t g_astc_tables[sizeof(AstcShared) / 4];
template <typename X>
__global__ void __launch_bounds__(256) k_uastc_blocks_16(const Ktx2File *files, int32_t *status2, const uint8_t *blob, uint8_t *O, const uint32_t *layer_list, const uint32_t *x_tables) {
    __shared__ __align__(16) UastcShared T; __shared__ __align__(16) X TX; __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {          // both table images by bulk asynchronous copies, one barrier
        mbar_init(&bar, 1); mbar_expect_tx(&bar, (uint32_t)(sizeof(UastcShared) + sizeof(X)));
        bulk_g2s(&T, g_tables, (uint32_t)sizeof(UastcShared), &bar); bulk_g2s(&TX, x_tables, (uint32_t)sizeof(X), &bar);
    }
    __shared__ struct { const uint8_t *src0; uint8_t *dst; uint32_t nblk, fi, skip; } K;
    if (threadIdx.x == 0) {
        const uint32_t ll = layer_list[blockIdx.y], fi = ll >> 12, L = ll & 4095;
        const Ktx2File &f = files[fi];
        K.skip = f.status || !f.is_uastc; K.fi = fi; K.nblk = f.bx * f.by;
        K.src0 = blob + f.file_off + f.level_off + (size_t)L * K.nblk * 16;
        K.dst = O + f.o_rgba + (size_t)L * K.nblk * 16;
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    if (K.skip) return;
    const uint32_t nblk = K.nblk, fi = K.fi;
    const uint8_t *src0 = K.src0; uint8_t *dst = K.dst;
    const bool aligned = (((uintptr_t)src0) & 15) == 0;
#pragma unroll 1
    for (uint32_t k = 0; k < UASTC_CHUNKS; k++) {
        const uint32_t bi = (blockIdx.x * UASTC_CHUNKS + k) * 256 + threadIdx.x;
        if (bi >= nblk) return;
        const uint8_t *src = src0 + (size_t)bi * 16;
        uint4 blk;
        if (aligned) blk = __ldcs((const uint4 *)src);
        else { uint32_t w[4]; for (int q = 0; q < 4; q++) w[q] = src[4 * q] | (src[4 * q + 1] << 8) | (src[4 * q + 2] << 16) | ((uint32_t)src[4 * q + 3] << 24); blk = make_uint4(w[0], w[1], w[2], w[3]); }
        uint32_t o[4]; bool ok;
        if constexpr (sizeof(X) == sizeof(Bc7Shared)) ok = uastc_to_bc7(T, TX, blk.x, blk.y, blk.z, blk.w, o);
        else ok = uastc_to_astc(T, TX, blk.x, blk.y, blk.z, blk.w, o);
        if (!ok) { status2[2 * fi] = UVOL_ERR_CORRUPT; continue; }
        __stcs((uint4 *)(dst + (size_t)bi * 16), make_uint4(o[0], o[1], o[2], o[3]));
    }
}
static_assert(sizeof(Bc7Shared) != sizeof(AstcShared), "k_uastc_blocks_16 tells its targets apart by the table image");

bool g_tables_ready[16] = {};
}  // namespace

__device__ __align__(16) uint32_t g_bc7_tables_etc1s[sizeof(Bc7Shared) / 4];      // second image for the ETC1S kernel in basis_transcode.cu (separate translation unit)
const uint32_t *uvol_bc7_tables_device() { uint32_t *p = nullptr; cudaGetSymbolAddress((void **)&p, g_bc7_tables_etc1s); return p; }
__device__ __align__(16) uint32_t g_eac_map[128];                                  // ETC1S alpha -> EAC parameters (basis_core.h etc1s_alpha_to_eac)
const uint32_t *uvol_eac_map_device() { uint32_t *p = nullptr; cudaGetSymbolAddress((void **)&p, g_eac_map); return p; }

// status2: the launcher's per-file {status, aux} pairs.  layer list entries: file << 12 | layer.
// Uploads the table images once per device (synchronous).
int uvol_texture_tables_ready(int device) {
    static std::mutex table_mu;
    std::lock_guard<std::mutex> table_lock(table_mu);
    if (device >= 0 && device < 16 && g_tables_ready[device]) return 0;
    UastcShared h; uastc_fill_tables(h);
    cudaError_t e = cudaMemcpyToSymbol(g_tables, &h, sizeof h);
    if (e != cudaSuccess) return (int)e;
    Bc7Shared b; bc7_fill_tables(b);
    e = cudaMemcpyToSymbol(g_bc7_tables, &b, sizeof b); if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_bc7_tables_etc1s, &b, sizeof b); if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_eac_map, ETC1S_EAC_MAP_INIT, sizeof ETC1S_EAC_MAP_INIT); if (e != cudaSuccess) return (int)e;
    AstcShared a; astc_fill_tables(a);
    e = cudaMemcpyToSymbol(g_astc_tables, &a, sizeof a); if (e != cudaSuccess) return (int)e;
    if (device >= 0 && device < 16) g_tables_ready[device] = true;
    return 0;
}

int uvol_uastc_launch(int device, const Ktx2File *dF, int32_t *status2, const uint8_t *dBlob, uint8_t *dOut, const uint32_t *dLayerList, int nlayers,
                      uint32_t max_blocks, int target, cudaStream_t st) {
    const int rc = uvol_texture_tables_ready(device); if (rc) return rc;
    const dim3 grid((max_blocks + 256 * UASTC_CHUNKS - 1) / (256 * UASTC_CHUNKS), (unsigned)nlayers);
    if (target == UVOL_TEX_BC7 || target == UVOL_TEX_ASTC_4x4) {
        uint32_t *xt = nullptr;
        if (target == UVOL_TEX_BC7) { cudaGetSymbolAddress((void **)&xt, g_bc7_tables); k_uastc_blocks_16<Bc7Shared><<<grid, 256, 0, st>>>(dF, status2, dBlob, dOut, dLayerList, xt); }
        else { cudaGetSymbolAddress((void **)&xt, g_astc_tables); k_uastc_blocks_16<AstcShared><<<grid, 256, 0, st>>>(dF, status2, dBlob, dOut, dLayerList, xt); }
        return 0;
    }
    static const bool no_tma = getenv("UVOL_NO_TMA") != nullptr;
    if (no_tma) { k_uastc_blocks<false><<<grid, 256, 0, st>>>(dF, status2, dBlob, dOut, dLayerList); return 0; }
    k_uastc_blocks<true><<<dim3((max_blocks + 256 * UASTC_CHUNKS - 1) / (256 * UASTC_CHUNKS), (unsigned)nlayers), 256, 0, st>>>(dF, status2, dBlob, dOut, dLayerList);
    return 0;
}
