// basis_transcode.cu -- sm_100a kernels + batch launcher for the V2 texture path (KTX2 -> RGBA32).
//
// Replaces KTX2Loader._createTexture -> BasisWorker.transcode (src/lib/KTX2Loader.js:297-337,
// 469-580) for a batch of .ktx2 segments.  Stages (DESIGN.md "Texture pipeline"):
//   globals   one warp per file : Huffman tables + endpoint/selector codebooks (startTranscoding, :506)
//   slices    one warp per slice: VLC decode -> per-block {pred, delta symbol, selector}   (transcodeImage, :551)
//   resolve   one warp per (file, plane): layer-ordered, row-ordered endpoint prediction reversal
//             (left / up / CR / delta) as a warp segmented scan; CR selectors copied from the previous layer
//   blocks    one thread per 4x4 block: codebook lookup -> 64 B of RGBA32, coalesced row stores
// UASTC files skip the first three stages (no entropy coding) and run the block kernel in uastc_transcode.cu.
#include <chrono>
#include <string.h>
#include <atomic>
#include <thread>
#include "uvol_ctx.h"
#include "basis_core.h"
#include "bc7_core.h"
#include "tma_bulk.h"

int uvol_ktx2_parse(const uint8_t *b, size_t len, uint32_t file_index, Ktx2File &f, std::vector<Ktx2Slice> &slices);
int uvol_ktx2_split_levels(const uint8_t *b, size_t len, std::vector<std::vector<uint8_t>> &out);
extern "C" int uvol_zstd_inflate(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len);
int uvol_uastc_launch(int device, const Ktx2File *dF, int32_t *status2, const uint8_t *dBlob, uint8_t *dOut, const uint32_t *dLayerList, int nlayers,
                      uint32_t max_blocks, int target, cudaStream_t st);
int uvol_texture_tables_ready(int device);
const uint32_t *uvol_bc7_tables_device();
const uint32_t *uvol_eac_map_device();

namespace {

struct TexState { int32_t status; uint32_t hist_size; };

#ifndef SERIAL_WARPS
#define SERIAL_WARPS 1        // one unit per warp, one warp per block (small blocks pack around the geometry stream's blocks)
#endif
__global__ void __launch_bounds__(32 * SERIAL_WARPS) k_basis_globals(const Ktx2File *files, TexState *state, const uint8_t *blob, uint8_t *S, int file0, int nfiles) {
    const uint32_t fi = (uint32_t)file0 + blockIdx.x * SERIAL_WARPS + (threadIdx.x >> 5);
    if ((int)fi >= nfiles || (threadIdx.x & 31) != 0) return;
    const Ktx2File &f = files[fi];
    if (f.status) { state[fi].status = f.status; return; }
    if (f.is_uastc) return;
    BasisGlobalsMem m;
    m.endpoints = (uint32_t *)(S + f.o_endpoints); m.selectors = (uint32_t *)(S + f.o_selectors); m.tables = (HuffTable *)(S + f.o_huff);
    m.pool = (uint16_t *)(S + f.o_sorted); m.pool_cap = f.endpoint_count + f.selector_count + 8192 + 1024;
    m.sizes = (uint8_t *)(m.pool + m.pool_cap);
    uint32_t hs = 0;
    const int rc = basis_build_globals(f, blob + f.file_off, m, &hs);
    state[fi].hist_size = hs;
    if (rc) state[fi].status = rc;
}

#define SLICE_SMEM_BYTES (4 * sizeof(HuffTable) + 4096 + 2048)
__global__ void __launch_bounds__(32 * SERIAL_WARPS) k_etc1s_slices(const Ktx2File *files, TexState *state, const Ktx2Slice *slices, const uint8_t *blob, uint8_t *S, int slice0, int nslices) {
    extern __shared__ uint4 slice_smem[];
    uint8_t *my = (uint8_t *)slice_smem + (size_t)(threadIdx.x >> 5) * SLICE_SMEM_BYTES;
    HuffTable *tabs = (HuffTable *)my; uint8_t *rowp = my + 4 * sizeof(HuffTable); uint16_t *hist = (uint16_t *)(rowp + 4096);
    const uint32_t si = (uint32_t)slice0 + blockIdx.x * SERIAL_WARPS + (threadIdx.x >> 5);
    if ((int)si >= nslices) return;
    const Ktx2Slice &sl = slices[si]; const Ktx2File &f = files[sl.file];
    if (f.status || state[sl.file].status) return;
    const HuffTable *gt = (const HuffTable *)(S + f.o_huff);
    {   // stage the four slice tables in shared memory (word copy by the whole warp)
        const uint32_t *src = (const uint32_t *)gt; uint32_t *dst = (uint32_t *)tabs;
        for (uint32_t i = threadIdx.x & 31; i < 4 * sizeof(HuffTable) / 4; i += 32) dst[i] = src[i];
    }
    __syncwarp();
    if ((threadIdx.x & 31) != 0) return;
    BitRd b; br_init(b, blob + f.file_off + sl.data_off, sl.data_len);
    SliceTables T{&tabs[0], &tabs[1], &tabs[2], &tabs[3], (const uint16_t *)(S + f.o_sorted)};
    const int rc = etc1s_slice_symbols(b, T, f.bx, f.by, f.selector_count, state[sl.file].hist_size, (int)f.is_video, rowp, hist,
                                       S + sl.o_pred, (uint16_t *)(S + sl.o_delta), (uint16_t *)(S + sl.o_sel));
    if (rc) state[sl.file].status = rc;
    else if ((b.consumed + 7) / 8 > sl.data_len) state[sl.file].status = UVOL_ERR_TRUNCATED;
}

// Endpoint prediction reversal.  One warp per (file, plane); layers and rows in order, 32 blocks per step.
__global__ void __launch_bounds__(32 * SERIAL_WARPS) k_etc1s_resolve(const Ktx2File *files, TexState *state, const Ktx2Slice *slices, uint8_t *S, int file0, int nfiles) {
    const uint32_t fi = (uint32_t)file0 + blockIdx.x * SERIAL_WARPS + (threadIdx.x >> 5), plane = blockIdx.y, lane = threadIdx.x & 31;
    if ((int)fi >= nfiles) return;
    const Ktx2File &f = files[fi];
    if (f.status || state[fi].status || f.is_uastc) return;
    if (plane && !f.has_alpha) return;
    const uint32_t bx = f.bx, by = f.by, ec = f.endpoint_count; const int video = (int)f.is_video;
    int bad = 0;
    for (uint32_t L = 0; L < f.layers; L++) {
        const Ktx2Slice &sl = slices[f.first_slice + plane * f.layers + L];
        const uint8_t *pred = S + sl.o_pred; const uint16_t *delta = (const uint16_t *)(S + sl.o_delta);
        uint16_t *E = (uint16_t *)(S + sl.o_ep), *Ssel = (uint16_t *)(S + sl.o_sel);
        const uint16_t *PE = nullptr, *PS = nullptr;
        if (L) { const Ktx2Slice &pl = slices[f.first_slice + plane * f.layers + L - 1]; PE = (const uint16_t *)(S + pl.o_ep); PS = (const uint16_t *)(S + pl.o_sel); }
        uint32_t carry = 0;
        for (uint32_t y = 0; y < by; y++) {
            for (uint32_t base = 0; base < bx; base += 32) {
                const uint32_t x = base + lane; const bool in = x < bx; const uint32_t bi = y * bx + x;
                uint32_t p = 0, d = 0, bv = 0; bool head = false;
                if (in) {
                    p = pred[bi];
                    if (p == 3) d = delta[bi];
                    else if (p == 1) { head = true; bv = E[bi - bx]; }
                    else if (p == 2) {
                        head = true;
                        if (video) { if (PE) { bv = PE[bi]; Ssel[bi] = PS[bi]; } else bad = 1; }
                        else bv = E[bi - bx - 1];
                    }
                }
                uint32_t sc = d;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, sc, o); if ((int)lane >= o) sc += t; }
                const unsigned hm = __ballot_sync(0xffffffffu, head);
                const unsigned below = hm & (0xffffffffu >> (31 - lane));
                const int h = below ? 31 - __clz(below) : 0;
                const uint32_t bh = __shfl_sync(0xffffffffu, bv, h), sh = __shfl_sync(0xffffffffu, sc, h);
                uint32_t e = below ? bh + sc - sh : carry + sc;
                e %= ec;
                if (in) E[bi] = (uint16_t)e;
                const int lastl = (bx - base) >= 32 ? 31 : (int)(bx - base) - 1;
                carry = __shfl_sync(0xffffffffu, e, lastl);
            }
            __syncwarp();
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) state[fi].status = UVOL_ERR_CORRUPT;
}

// Block -> RGBA32.  grid = (ceil(nblk / (256 * ETC1S_CHUNKS)), layer list); one thread per 4x4 block, each pixel row of
// a block is one 16-byte store, a warp covers 32 adjacent blocks = 512 contiguous bytes per row.  The per-layer constants
// (descriptor fields, array bases) are fetched once per CTA into shared memory: as per-thread global loads they filled the
// load/store queue (ncu: lg_throttle 19 of 43 stall cycles per instruction) ahead of the stores that matter.
// With SMEM_CB the file's endpoint / selector codebooks (4 B per entry) are staged in shared memory by the CTA: as global gathers
// the two random 4-byte lookups per block cost 32 L1 wavefronts each and saturate the L1 pipe (ncu: l1tex 89 %, DRAM 22 %), from
// shared memory a random lookup is a few bank-conflict cycles.  A CTA decodes ETC1S_CHUNKS x 256 blocks per staged codebook.
#define ETC1S_CHUNKS 16
struct BlockLayerConst { const uint32_t *eps, *sels; const uint16_t *ep_idx, *sel_idx, *aep_idx, *asel_idx; uint8_t *dst; uint32_t nblk, bx, W, H, has_alpha, skip, ec, sc; };
// The staging itself is two bulk asynchronous copies (cp.async.bulk into shared memory, completion counted on an mbarrier) issued by
// one thread -- the TMA unit moves the codebooks while the CTA's threads already fetch their blocks' indices; use_tma = 0 keeps the
// copy loop by all threads for the A/B measurement (UVOL_NO_TMA=1).
template <bool SMEM_CB>
__global__ void __launch_bounds__(256) k_etc1s_blocks(const Ktx2File *files, const TexState *state, const Ktx2Slice *slices, const uint32_t *layer_list,
                                                      const uint8_t *S, uint8_t *O, int use_tma) {
    __shared__ BlockLayerConst K; __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        const uint32_t ll = layer_list[blockIdx.y], fi = ll >> 12, L = ll & 4095;
        const Ktx2File &f = files[fi];
        K.skip = f.status || state[fi].status || f.is_uastc;
        if (!K.skip) {
            const Ktx2Slice &sl = slices[f.first_slice + L];
            K.eps = (const uint32_t *)(S + f.o_endpoints); K.sels = (const uint32_t *)(S + f.o_selectors);
            K.ep_idx = (const uint16_t *)(S + sl.o_ep); K.sel_idx = (const uint16_t *)(S + sl.o_sel);
            K.has_alpha = f.has_alpha;
            if (f.has_alpha) { const Ktx2Slice &al = slices[f.first_slice + f.layers + L]; K.aep_idx = (const uint16_t *)(S + al.o_ep); K.asel_idx = (const uint16_t *)(S + al.o_sel); }
            K.nblk = f.bx * f.by; K.bx = f.bx; K.W = f.width; K.H = f.height; K.ec = f.endpoint_count; K.sc = f.selector_count;
            K.dst = O + f.o_rgba + (size_t)L * f.width * f.height * 4;
        }
    }
    __syncthreads();
    if (K.skip) return;
    extern __shared__ __align__(16) uint32_t cb_smem[];
    const uint32_t *eps = K.eps, *sels = K.sels;
    if (SMEM_CB) {
        if (blockIdx.x * (ETC1S_CHUNKS * 256u) >= K.nblk) return;
        const uint32_t ec4 = (K.ec + 3u) & ~3u, sc4 = (K.sc + 3u) & ~3u;          // 16-byte granules; the arena pads every array to 128 bytes, so the tail is readable
        if (use_tma) {
            if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_expect_tx(&bar, (ec4 + sc4) * 4u); bulk_g2s(cb_smem, K.eps, ec4 * 4u, &bar); bulk_g2s(cb_smem + ec4, K.sels, sc4 * 4u, &bar); }
            __syncthreads();
            mbar_wait(&bar, 0);
        } else {
            for (uint32_t i = threadIdx.x; i < K.ec; i += 256) cb_smem[i] = K.eps[i];
            for (uint32_t i = threadIdx.x; i < K.sc; i += 256) cb_smem[ec4 + i] = K.sels[i];
            __syncthreads();
        }
        eps = cb_smem; sels = cb_smem + ec4;
    }
    const uint32_t nblk = K.nblk, bxn = K.bx, W = K.W, H = K.H;
    const bool whole = (W & 3) == 0;
#pragma unroll 1
    for (uint32_t k = 0; k < ETC1S_CHUNKS; k++) {
        const uint32_t bi = (blockIdx.x * ETC1S_CHUNKS + k) * 256 + threadIdx.x;
        if (bi >= nblk) return;
        const uint32_t ecm = K.ec - 1, scm = K.sc - 1;                      // (indices are validated upstream; the clamp only keeps a corrupt file inside the staged tables)
        const uint32_t ep = eps[min((uint32_t)K.ep_idx[bi], ecm)], se = sels[min((uint32_t)K.sel_idx[bi], scm)];
        uint32_t rows[4][4];
        etc1s_block_rows(ep, se, rows);
        if (K.has_alpha) {
            const uint32_t aep = eps[min((uint32_t)K.aep_idx[bi], ecm)], ase = sels[min((uint32_t)K.asel_idx[bi], scm)];
            const uint32_t a0 = (etc1s_color(aep, 0) >> 8) & 255u, a1 = (etc1s_color(aep, 1) >> 8) & 255u, a2 = (etc1s_color(aep, 2) >> 8) & 255u, a3 = (etc1s_color(aep, 3) >> 8) & 255u;      // alpha = G of the alpha slice
#pragma unroll
            for (int y = 0; y < 4; y++) {
                const uint32_t rb = (ase >> (8 * y)) & 255u;
#pragma unroll
                for (int x = 0; x < 4; x++) { const uint32_t q = (rb >> (2 * x)) & 3u; rows[y][x] = (rows[y][x] & 0x00ffffffu) | (((q & 2u) ? ((q & 1u) ? a3 : a2) : ((q & 1u) ? a1 : a0)) << 24); }
            }
        }
        const uint32_t xb = bi % bxn, yb = bi / bxn;
        uint8_t *dst = K.dst;
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

// Block -> ETC1 (target format ETC1): 4 B of indices in, one 8-byte block out per thread; a warp writes 256 contiguous bytes.
// grid = (ceil(nblk / 256), layer list).  Opaque ETC1S files only (an alpha slice would need an EAC block next to it).
__global__ void __launch_bounds__(256) k_etc1s_blocks_etc1(const Ktx2File *files, const TexState *state, const Ktx2Slice *slices, const uint32_t *layer_list,
                                                           const uint8_t *S, uint8_t *O) {
    const uint32_t ll = layer_list[blockIdx.y], fi = ll >> 12, L = ll & 4095;
    const Ktx2File &f = files[fi];
    if (f.status || state[fi].status || f.is_uastc || f.has_alpha) return;
    const uint32_t nblk = f.bx * f.by, bi = blockIdx.x * 256 + threadIdx.x;
    if (bi >= nblk) return;
    const Ktx2Slice &sl = slices[f.first_slice + L];
    const uint32_t *eps = (const uint32_t *)(S + f.o_endpoints), *sels = (const uint32_t *)(S + f.o_selectors);
    const uint32_t ei = min((uint32_t)((const uint16_t *)(S + sl.o_ep))[bi], f.endpoint_count - 1), si = min((uint32_t)((const uint16_t *)(S + sl.o_sel))[bi], f.selector_count - 1);
    const Etc1Words w = etc1s_to_etc1(eps[ei], sels[si]);
    ((uint2 *)(O + f.o_rgba + (size_t)L * nblk * 8))[bi] = make_uint2(w.x, w.y);
}

// Block -> ETC2 RGBA (target format ETC2_RGBA: `etc2Supported`, src/lib/KTX2Loader.js:619-627): 4 B (8 B with an alpha slice) of indices in,
// one 16-byte block out per thread -- the EAC alpha block (basis_core.h etc1s_alpha_to_eac; constant 255 for opaque files), then the
// ETC1 colour block.  grid = (ceil(nblk / 256), layer list).
__global__ void __launch_bounds__(256) k_etc1s_blocks_etc2a(const Ktx2File *files, const TexState *state, const Ktx2Slice *slices, const uint32_t *layer_list,
                                                            const uint8_t *S, uint8_t *O, const uint32_t *eac_map) {
    __shared__ uint32_t map[128];
    if (threadIdx.x < 128) map[threadIdx.x] = eac_map[threadIdx.x];
    __syncthreads();
    const uint32_t ll = layer_list[blockIdx.y], fi = ll >> 12, L = ll & 4095;
    const Ktx2File &f = files[fi];
    if (f.status || state[fi].status || f.is_uastc) return;
    const uint32_t nblk = f.bx * f.by, bi = blockIdx.x * 256 + threadIdx.x;
    if (bi >= nblk) return;
    const Ktx2Slice &sl = slices[f.first_slice + L];
    const uint32_t *eps = (const uint32_t *)(S + f.o_endpoints), *sels = (const uint32_t *)(S + f.o_selectors);
    const uint32_t ecm = f.endpoint_count - 1, scm = f.selector_count - 1;
    const uint32_t ei = min((uint32_t)((const uint16_t *)(S + sl.o_ep))[bi], ecm), si = min((uint32_t)((const uint16_t *)(S + sl.o_sel))[bi], scm);
    const Etc1Words w = etc1s_to_etc1(eps[ei], sels[si]);
    EacWords a = eac_opaque();
    if (f.has_alpha) {
        const Ktx2Slice &al = slices[f.first_slice + f.layers + L];
        a = etc1s_alpha_to_eac(eps[min((uint32_t)((const uint16_t *)(S + al.o_ep))[bi], ecm)], sels[min((uint32_t)((const uint16_t *)(S + al.o_sel))[bi], scm)], map);
    }
    __stcs((uint4 *)(O + f.o_rgba + (size_t)L * nblk * 16) + bi, make_uint4(a.x, a.y, w.x, w.y));
}

// Block -> BC1 / BC3 (`dxtSupported`, src/lib/KTX2Loader.js:610-618): one 8-byte BC1 block, or the BC4 alpha block + the BC1 block, per thread
// (basis_core.h etc1s_to_bc1 / etc1s_alpha_to_bc4).  grid = (ceil(nblk / 256), layer list).
__global__ void __launch_bounds__(256) k_etc1s_blocks_dxt(const Ktx2File *files, const TexState *state, const Ktx2Slice *slices, const uint32_t *layer_list,
                                                          const uint8_t *S, uint8_t *O, int bc3) {
    const uint32_t ll = layer_list[blockIdx.y], fi = ll >> 12, L = ll & 4095;
    const Ktx2File &f = files[fi];
    if (f.status || state[fi].status || f.is_uastc) return;
    const uint32_t nblk = f.bx * f.by, bi = blockIdx.x * 256 + threadIdx.x;
    if (bi >= nblk) return;
    const Ktx2Slice &sl = slices[f.first_slice + L];
    const uint32_t *eps = (const uint32_t *)(S + f.o_endpoints), *sels = (const uint32_t *)(S + f.o_selectors);
    const uint32_t ecm = f.endpoint_count - 1, scm = f.selector_count - 1;
    const uint32_t ei = min((uint32_t)((const uint16_t *)(S + sl.o_ep))[bi], ecm), si = min((uint32_t)((const uint16_t *)(S + sl.o_sel))[bi], scm);
    const Bc1Words w = etc1s_to_bc1(eps[ei], sels[si]);
    if (!bc3) { ((uint2 *)(O + f.o_rgba + (size_t)L * nblk * 8))[bi] = make_uint2(w.x, w.y); return; }
    Bc1Words a = bc4_opaque();
    if (f.has_alpha) {
        const Ktx2Slice &al = slices[f.first_slice + f.layers + L];
        a = etc1s_alpha_to_bc4(eps[min((uint32_t)((const uint16_t *)(S + al.o_ep))[bi], ecm)], sels[min((uint32_t)((const uint16_t *)(S + al.o_sel))[bi], scm)]);
    }
    __stcs((uint4 *)(O + f.o_rgba + (size_t)L * nblk * 16) + bi, make_uint4(a.x, a.y, w.x, w.y));
}

// Block -> BC7 mode 5 (target format BC7, src/lib/KTX2Loader.js:602-604): 4 B (8 B with an alpha slice) of indices in, one 16-byte block
// out per thread; a warp writes 512 contiguous bytes.  grid = (ceil(nblk / 256), layer list); per-block logic in bc7_core.h.
__global__ void __launch_bounds__(256) k_etc1s_blocks_bc7(const Ktx2File *files, const TexState *state, const Ktx2Slice *slices, const uint32_t *layer_list,
                                                          const uint8_t *S, uint8_t *O, const uint32_t *bc7_tables) {
    __shared__ Bc7Shared B7;
    for (uint32_t i = threadIdx.x; i < sizeof(Bc7Shared) / 4; i += 256) ((uint32_t *)&B7)[i] = bc7_tables[i];
    __syncthreads();
    const uint32_t ll = layer_list[blockIdx.y], fi = ll >> 12, L = ll & 4095;
    const Ktx2File &f = files[fi];
    if (f.status || state[fi].status || f.is_uastc) return;
    const uint32_t nblk = f.bx * f.by, bi = blockIdx.x * 256 + threadIdx.x;
    if (bi >= nblk) return;
    const Ktx2Slice &sl = slices[f.first_slice + L];
    const uint32_t *eps = (const uint32_t *)(S + f.o_endpoints), *sels = (const uint32_t *)(S + f.o_selectors);
    const uint32_t ecm = f.endpoint_count - 1, scm = f.selector_count - 1;
    const uint32_t ei = min((uint32_t)((const uint16_t *)(S + sl.o_ep))[bi], ecm), si = min((uint32_t)((const uint16_t *)(S + sl.o_sel))[bi], scm);
    uint32_t aep = 0, asel = 0;
    if (f.has_alpha) {
        const Ktx2Slice &al = slices[f.first_slice + f.layers + L];
        aep = eps[min((uint32_t)((const uint16_t *)(S + al.o_ep))[bi], ecm)]; asel = sels[min((uint32_t)((const uint16_t *)(S + al.o_sel))[bi], scm)];
    }
    uint32_t o[4];
    etc1s_to_bc7(B7, eps[ei], sels[si], f.has_alpha != 0, aep, asel, o);
    __stcs((uint4 *)(O + f.o_rgba + (size_t)L * nblk * 16) + bi, make_uint4(o[0], o[1], o[2], o[3]));
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
uint64_t take(uint64_t &cur, uint64_t bytes) { uint64_t o = cur; cur = (cur + bytes + 127) / 128 * 128; return o; }

}  // namespace

static const char *kTexStages[] = {"h2d", "globals", "slices", "resolve", "blocks", "d2h"};
extern "C" const char *uvol_tex_stage_name(int i) { return (i >= 0 && i < 6) ? kTexStages[i] : ""; }

struct TexBatch {
    std::vector<Ktx2File> files; std::vector<Ktx2Slice> slices; std::vector<uint32_t> layer_list, uastc_layers;
    int n = 0; uint64_t blob_bytes = 0, scratch = 0, out = 0, bytes_in = 0; uint32_t max_blocks = 1, max_codebook = 0; bool any_alpha = false, any_zstd = false; int target = UVOL_TEX_RGBA32;
    size_t desc_bytes = 0, off_sl = 0, off_ll = 0, off_ul = 0; double parse_ms = 0; uint32_t launches = 0; int nev = 0;
    std::vector<uint32_t> ll_start, ul_start, sl_start;      // per file (+1): first entry in layer_list / uastc_layers / slices
    // mip chains: the caller's files [0, n_user) -> the single-level files [0, n) the machinery works on (uvol_ktx2_split_levels)
    int n_user = 0; std::vector<uint32_t> first_of, nlev; std::vector<int32_t> split_status; std::vector<std::vector<uint8_t>> synth;
    std::vector<const uint8_t *> xdata; std::vector<size_t> xsize; std::vector<uvol_texture_level> mips;
};
void uvol_tex_batch_free(TexBatch *b) { delete b; }

// Container parse of every file, memory plan, descriptor upload.  The payload bytes are NOT staged here (see ktx2_stage_range).
static int ktx2_prepare(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int target = UVOL_TEX_RGBA32) {
    if (!ctx->tex) ctx->tex = new TexBatch();
    TexBatch &B = *ctx->tex;
    B.n = n; B.files.assign((size_t)n, Ktx2File()); B.slices.clear(); B.layer_list.clear(); B.uastc_layers.clear();
    B.ll_start.assign((size_t)n + 1, 0); B.ul_start.assign((size_t)n + 1, 0); B.sl_start.assign((size_t)n + 1, 0);
    B.max_blocks = 1; B.max_codebook = 0; B.any_alpha = false; B.any_zstd = false; B.bytes_in = 0; B.target = target;
    std::vector<Ktx2File> &files = B.files; std::vector<Ktx2Slice> &slices = B.slices;
    uint64_t blob_bytes = 0, s = 0, o = 0;
    for (int i = 0; i < n; i++) {
        Ktx2File &f = files[i]; memset(&f, 0, sizeof f);
        f.file_len = (uint32_t)size[i]; B.bytes_in += size[i];
        B.ll_start[i] = (uint32_t)B.layer_list.size(); B.ul_start[i] = (uint32_t)B.uastc_layers.size(); B.sl_start[i] = (uint32_t)slices.size();
        const size_t slices_before = slices.size();
        f.status = (data[i] && size[i] < (1ull << 31)) ? uvol_ktx2_parse(data[i], size[i], (uint32_t)i, f, slices) : UVOL_ERR_ARG;
        f.file_off = blob_bytes;                                      // the file itself, or (Zstd levels) the inflated level
        blob_bytes = align_up(blob_bytes + (!f.status && f.zstd ? std::max<uint64_t>(f.z_len, size[i]) : (uint64_t)size[i]) + 8, 16);   // (a file rejected later is still copied as is)
        if (!f.status && f.zstd) B.any_zstd = true;
        for (size_t k = slices_before; k < slices.size() && !f.status; k++)      // what the slice kernel dereferences lies inside this file (the parser guarantees it; checked again at the trust boundary)
            if ((uint64_t)slices[k].data_off + slices[k].data_len > size[i]) f.status = UVOL_ERR_CORRUPT;
        if (f.status) { slices.resize(slices_before); continue; }
        if (f.layers > 4095 || f.bx > 4096) { f.status = UVOL_ERR_UNSUPPORTED; slices.resize(slices_before); continue; }
        const uint64_t nblk = (uint64_t)f.bx * f.by;
        if (target == UVOL_TEX_ETC1 && (f.is_uastc || f.has_alpha)) { f.status = UVOL_ERR_UNSUPPORTED; slices.resize(slices_before); continue; }   // ETC1 target: opaque ETC1S sources only
        if ((target == UVOL_TEX_BC1 || target == UVOL_TEX_BC3) && f.is_uastc) { f.status = UVOL_ERR_UNSUPPORTED; slices.resize(slices_before); continue; }   // BC1 / BC3 targets: ETC1S sources only
        if (target == UVOL_TEX_ETC2_RGBA && f.is_uastc) { f.status = UVOL_ERR_UNSUPPORTED; slices.resize(slices_before); continue; }            // ETC2 RGBA target: ETC1S sources only (UASTC would need an ETC1 encoder)
        if (target == UVOL_TEX_ASTC_4x4 && !f.is_uastc) { f.status = UVOL_ERR_UNSUPPORTED; slices.resize(slices_before); continue; }             // ASTC target: UASTC sources only (KTX2Loader.js:592-600)
        const uint64_t out_bytes = target == UVOL_TEX_ETC1 || target == UVOL_TEX_BC1 ? (uint64_t)f.layers * nblk * 8 : (target == UVOL_TEX_BC7 || target == UVOL_TEX_ASTC_4x4 || target == UVOL_TEX_ETC2_RGBA || target == UVOL_TEX_BC3 ? (uint64_t)f.layers * nblk * 16 : (uint64_t)f.layers * f.width * f.height * 4);
        if (out_bytes > ctx->cfg.max_texture_bytes) { f.status = UVOL_ERR_UNSUPPORTED; slices.resize(slices_before); continue; }      // resource limit, per item
        if (nblk > B.max_blocks) B.max_blocks = (uint32_t)nblk;
        f.o_rgba = take(o, out_bytes);
        if (f.is_uastc) { for (uint32_t L = 0; L < f.layers; L++) B.uastc_layers.push_back(((uint32_t)i << 12) | L); continue; }   // no entropy stage, no scratch
        B.any_alpha |= f.has_alpha != 0;
        if (f.endpoint_count + f.selector_count > B.max_codebook) B.max_codebook = f.endpoint_count + f.selector_count;
        const uint64_t pool = (uint64_t)f.endpoint_count + f.selector_count + 8192 + 1024;
        f.o_endpoints = take(s, (uint64_t)f.endpoint_count * 4); f.o_selectors = take(s, (uint64_t)f.selector_count * 4);
        f.o_huff = take(s, sizeof(HuffTable) * 10); f.o_sorted = take(s, pool * 2 + 32768 + 64);
        for (size_t k = f.first_slice; k < slices.size(); k++) {
            Ktx2Slice &sl = slices[k];
            sl.o_pred = take(s, nblk); sl.o_delta = take(s, nblk * 2); sl.o_sel = take(s, nblk * 2); sl.o_ep = take(s, nblk * 2);
        }
        for (uint32_t L = 0; L < f.layers; L++) B.layer_list.push_back(((uint32_t)i << 12) | L);
    }
    B.ll_start[n] = (uint32_t)B.layer_list.size(); B.ul_start[n] = (uint32_t)B.uastc_layers.size(); B.sl_start[n] = (uint32_t)slices.size();
    B.blob_bytes = blob_bytes; B.scratch = s; B.out = o;
    const size_t nsl = slices.size(), nll = B.layer_list.size(), nul = B.uastc_layers.size();
    B.off_sl = sizeof(Ktx2File) * (size_t)n; B.off_ll = B.off_sl + sizeof(Ktx2Slice) * (nsl + 1); B.off_ul = B.off_ll + 4 * (nll + 1);
    B.desc_bytes = B.off_ul + 4 * (nul + 1);
    UVOL_CUDA(ctx, ctx->h_tblob.reserve(blob_bytes + 64));
    UVOL_CUDA(ctx, ctx->h_tdesc.reserve(B.desc_bytes));
    UVOL_CUDA(ctx, ctx->d_tdesc.reserve(B.desc_bytes));
    UVOL_CUDA(ctx, ctx->d_tblob.reserve(blob_bytes + 64));
    UVOL_CUDA(ctx, ctx->d_tslices.reserve(sizeof(TexState) * (size_t)n));
    UVOL_CUDA(ctx, ctx->d_tscratch.reserve(s + 256));
    UVOL_CUDA(ctx, ctx->d_out_tex.reserve(o + 256));
    return UVOL_OK;
}

// Stages the payload of files [i0, i1) into the pinned blob with a few host threads: a copy of the file, or the inflated level of a
// Zstd-supercompressed file.  (A file whose inflate fails is marked; its descriptor must be uploaded after this.)
static void ktx2_stage_range(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int i0, int i1) {
    TexBatch &B = *ctx->tex; std::vector<Ktx2File> &files = B.files;
    std::atomic<int> next{i0}; uint64_t bytes = 0;
    for (int i = i0; i < i1; i++) bytes += size[i];
    auto work = [&]() {
        for (int i; (i = next.fetch_add(1)) < i1;) {
            if (!data[i] || size[i] >= (1ull << 31)) continue;
            uint8_t *dst = (uint8_t *)ctx->h_tblob.p + files[i].file_off;
            if (!files[i].status && files[i].zstd) {
                size_t got = 0;
                const int rc = uvol_zstd_inflate(data[i] + files[i].z_src_off, files[i].z_src_len, dst, files[i].z_len, &got);
                if (rc || got != files[i].z_len) files[i].status = rc ? rc : UVOL_ERR_CORRUPT;
            } else memcpy(dst, data[i], size[i]);
        }
    };
    int nthreads = bytes > (64ull << 20) || B.any_zstd ? (ctx->cfg.staging_threads ? (int)ctx->cfg.staging_threads : uvol_staging_threads()) : 1;
    if (B.any_zstd && !ctx->cfg.staging_threads && !getenv("UVOL_STAGING_THREADS")) nthreads = (int)std::min<unsigned>(32, std::max(1u, std::thread::hardware_concurrency()));   // inflating is ~20x slower than copying
    if (nthreads > i1 - i0) nthreads = i1 - i0;
    if (nthreads <= 1) work();
    else { std::vector<std::thread> pool; for (int t = 0; t < nthreads; t++) pool.emplace_back(work); for (auto &t : pool) t.join(); }
}

// Enqueues the kernels of files [i0, i1) on `st` (no host sync).  Every stage works on per-file / per-slice / per-layer lists, so a
// range of files is a range of each list.
static int ktx2_launch_range(uvol_ctx *ctx, int i0, int i1, cudaStream_t st, bool stamps, int *ev) {
    TexBatch &B = *ctx->tex; const int n = B.n;
    auto stamp = [&]() { if (stamps && ctx->profile && *ev < 8) cudaEventRecord(ctx->tex_ev[*ev], st); if (stamps) ++*ev; };
    const uint8_t *dD = (const uint8_t *)ctx->d_tdesc.p;
    const Ktx2File *dF = (const Ktx2File *)dD; const Ktx2Slice *dSl = (const Ktx2Slice *)(dD + B.off_sl);
    const uint32_t *dLL = (const uint32_t *)(dD + B.off_ll), *dUL = (const uint32_t *)(dD + B.off_ul);
    TexState *dSt = (TexState *)ctx->d_tslices.p; const uint8_t *dBlob = (const uint8_t *)ctx->d_tblob.p;
    uint8_t *dS = (uint8_t *)ctx->d_tscratch.p, *dO = (uint8_t *)ctx->d_out_tex.p;
    const uint32_t sl0 = B.sl_start[i0], sl1 = B.sl_start[i1], ll0 = B.ll_start[i0], ll1 = B.ll_start[i1], ul0 = B.ul_start[i0], ul1 = B.ul_start[i1];
    const unsigned nbf = (unsigned)((i1 - i0 + SERIAL_WARPS - 1) / SERIAL_WARPS);
    (void)n;
    k_basis_globals<<<nbf, 32 * SERIAL_WARPS, 0, st>>>(dF, dSt, dBlob, dS, i0, i1); B.launches++;
    stamp();
    if (sl1 > sl0) {
        cudaFuncSetAttribute(k_etc1s_slices, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SLICE_SMEM_BYTES * SERIAL_WARPS));
        k_etc1s_slices<<<(unsigned)((sl1 - sl0 + SERIAL_WARPS - 1) / SERIAL_WARPS), 32 * SERIAL_WARPS, SLICE_SMEM_BYTES * SERIAL_WARPS, st>>>(dF, dSt, dSl, dBlob, dS, (int)sl0, (int)sl1); B.launches++;
    }
    stamp();
    if (sl1 > sl0) { k_etc1s_resolve<<<dim3(nbf, B.any_alpha ? 2 : 1), 32 * SERIAL_WARPS, 0, st>>>(dF, dSt, dSl, dS, i0, i1); B.launches++; }
    stamp();
    const unsigned nll = ll1 - ll0, nul = ul1 - ul0;
    if (nll && B.target == UVOL_TEX_BC7) {
        UVOL_CUDA(ctx, (cudaError_t)uvol_texture_tables_ready(ctx->device));
        k_etc1s_blocks_bc7<<<dim3((B.max_blocks + 255) / 256, nll), 256, 0, st>>>(dF, dSt, dSl, dLL + ll0, dS, dO, uvol_bc7_tables_device()); B.launches++;
    } else if (nll && (B.target == UVOL_TEX_BC1 || B.target == UVOL_TEX_BC3)) {
        k_etc1s_blocks_dxt<<<dim3((B.max_blocks + 255) / 256, nll), 256, 0, st>>>(dF, dSt, dSl, dLL + ll0, dS, dO, B.target == UVOL_TEX_BC3 ? 1 : 0); B.launches++;
    } else if (nll && B.target == UVOL_TEX_ETC2_RGBA) {
        UVOL_CUDA(ctx, (cudaError_t)uvol_texture_tables_ready(ctx->device));
        k_etc1s_blocks_etc2a<<<dim3((B.max_blocks + 255) / 256, nll), 256, 0, st>>>(dF, dSt, dSl, dLL + ll0, dS, dO, uvol_eac_map_device()); B.launches++;
    } else if (nll && B.target == UVOL_TEX_ETC1) { k_etc1s_blocks_etc1<<<dim3((B.max_blocks + 255) / 256, nll), 256, 0, st>>>(dF, dSt, dSl, dLL + ll0, dS, dO); B.launches++; }
    else if (nll) {
        const dim3 grid((B.max_blocks + 256 * ETC1S_CHUNKS - 1) / (256 * ETC1S_CHUNKS), nll);
        const size_t cb = (size_t)B.max_codebook * 4 + 32;                    // (+ the 16-byte rounding of both codebooks)
        static const int use_tma = getenv("UVOL_NO_TMA") ? 0 : 1;
        if (cb <= 160 * 1024) {
            if (cb > 48 * 1024) UVOL_CUDA(ctx, cudaFuncSetAttribute(k_etc1s_blocks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cb));
            k_etc1s_blocks<true><<<grid, 256, cb, st>>>(dF, dSt, dSl, dLL + ll0, dS, dO, use_tma);
        } else k_etc1s_blocks<false><<<grid, 256, 0, st>>>(dF, dSt, dSl, dLL + ll0, dS, dO, 0);
        B.launches++;
    }
    if (nul) { UVOL_CUDA(ctx, (cudaError_t)uvol_uastc_launch(ctx->device, dF, (int32_t *)dSt, dBlob, dO, dUL + ul0, (int)nul, B.max_blocks, B.target, st)); B.launches++; }
    stamp();
    return UVOL_OK;
}

// Fresh batch, host buffers in: staging copy, upload, kernels and result copy are PIPELINED over chunks of whole files -- chunk k's
// kernels and result copy run while chunk k+1 is being staged and uploaded -- so the device -> host link starts carrying results a
// few tens of milliseconds into the call instead of after the whole batch has been staged, uploaded and transcoded.
static int ktx2_fresh(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int target, int memory, cudaStream_t st) {
    const double t_begin = now_ms();
    if (!ctx->tex) ctx->tex = new TexBatch();
    {   // files with a mip chain are taken apart into one single-level file per level; everything below sees only those
        TexBatch &X = *ctx->tex; X.n_user = n; X.first_of.assign((size_t)n, 0); X.nlev.assign((size_t)n, 1); X.split_status.assign((size_t)n, 0); X.synth.clear(); X.xdata.clear(); X.xsize.clear();
        bool any = false;
        for (int i = 0; i < n && !any; i++) any = data[i] && size[i] >= 44 && size[i] < (1ull << 31) && (data[i][40] | data[i][41] << 8 | data[i][42] << 16 | (uint32_t)data[i][43] << 24) > 1u;
        if (any) {
            std::vector<std::pair<int, int>> ref;          // per internal file: (caller's file, index into synth or -1)
            for (int i = 0; i < n; i++) {
                X.first_of[i] = (uint32_t)ref.size();
                const size_t before = X.synth.size();
                const int L = (data[i] && size[i] < (1ull << 31)) ? uvol_ktx2_split_levels(data[i], size[i], X.synth) : 0;
                if (L > 0) { X.nlev[i] = (uint32_t)L; for (int k = 0; k < L; k++) ref.push_back({i, (int)(before + k)}); }
                else { X.synth.resize(before); X.split_status[i] = L; ref.push_back({i, -1}); }
            }
            for (auto &r : ref) { if (r.second >= 0) { X.xdata.push_back(X.synth[r.second].data()); X.xsize.push_back(X.synth[r.second].size()); } else { X.xdata.push_back(data[r.first]); X.xsize.push_back(size[r.first]); } }
            if (X.xdata.size() >= (1u << 19)) { ctx->set_error("too many mip levels in one batch"); return UVOL_ERR_ARG; }
            data = X.xdata.data(); size = X.xsize.data(); n = (int)X.xdata.size();
        } else for (int i = 0; i < n; i++) X.first_of[i] = (uint32_t)i;
    }
    int rc = ktx2_prepare(ctx, data, size, n, target); if (rc) return rc;
    TexBatch &B = *ctx->tex; B.launches = 0;
    const size_t st_bytes = align_up(sizeof(TexState) * (size_t)n, 256);
    UVOL_CUDA(ctx, ctx->h_tstate.reserve(st_bytes));                       // per-ctx even when the bulk result buffer is shared
    if (memory == UVOL_MEM_HOST) UVOL_CUDA(ctx, ctx->ph_tout->reserve(B.out + 256));
    // chunks of about 1/8 of the input (at least 64 MB) -- a small first chunk gets the pipeline going.  Uploads and kernels run on
    // `st`, the result copies on s4 behind a per-chunk event, so chunk k's copy overlaps chunk k+1's upload and kernels.
    const uint64_t per = std::max<uint64_t>(64ull << 20, B.bytes_in / 8 + 1);
    int chunk = 0;
    int ev = 0; bool first = true;
    if (ctx->profile) cudaEventRecord(ctx->tex_ev[0], st);
    ev = 1;
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_tslices.p, 0, sizeof(TexState) * (size_t)n, st));
    uint8_t *hd = (uint8_t *)ctx->h_tdesc.p;
    const size_t nsl = B.slices.size(), nll = B.layer_list.size(), nul = B.uastc_layers.size();
    if (nsl) memcpy(hd + B.off_sl, B.slices.data(), sizeof(Ktx2Slice) * nsl);
    if (nll) memcpy(hd + B.off_ll, B.layer_list.data(), 4 * nll);
    if (nul) memcpy(hd + B.off_ul, B.uastc_layers.data(), 4 * nul);
    UVOL_CUDA(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_tdesc.p + B.off_sl, hd + B.off_sl, B.desc_bytes - B.off_sl, cudaMemcpyHostToDevice, st));
    uint8_t *hO = memory == UVOL_MEM_HOST ? (uint8_t *)ctx->ph_tout->p : nullptr, *dO = (uint8_t *)ctx->d_out_tex.p;
    for (int i0 = 0; i0 < n;) {
        int i1 = i0; uint64_t acc = 0;
        const uint64_t want = first ? std::min<uint64_t>(per, 96ull << 20) : per;
        while (i1 < n && (i1 == i0 || acc + size[i1] <= want)) acc += size[i1++];
        ktx2_stage_range(ctx, data, size, i0, i1);
        // descriptors of this chunk (a failed inflate has just been marked), then its bytes
        memcpy(hd + sizeof(Ktx2File) * (size_t)i0, B.files.data() + i0, sizeof(Ktx2File) * (size_t)(i1 - i0));
        UVOL_CUDA(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_tdesc.p + sizeof(Ktx2File) * (size_t)i0, hd + sizeof(Ktx2File) * (size_t)i0, sizeof(Ktx2File) * (size_t)(i1 - i0), cudaMemcpyHostToDevice, st));
        const uint64_t b0 = B.files[i0].file_off, b1 = i1 < n ? B.files[i1].file_off : B.blob_bytes;
        UVOL_CUDA(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_tblob.p + b0, (uint8_t *)ctx->h_tblob.p + b0, b1 - b0, cudaMemcpyHostToDevice, st));
        if (first) { if (ctx->profile && ev < 8) cudaEventRecord(ctx->tex_ev[ev], st); ev++; B.parse_ms = now_ms() - t_begin; }
        const bool last = i1 == n;
        rc = ktx2_launch_range(ctx, i0, i1, st, last, &ev); if (rc) return rc;      // (stage stamps on the last chunk: "blocks" then spans the pipelined middle)
        if (hO) {
            uint64_t o0 = ~0ull, o1 = 0;
            for (int i = i0; i < i1; i++) if (!B.files[i].status) { const uint64_t a = B.files[i].o_rgba; if (a < o0) o0 = a; }
            o1 = B.out; for (int i = i1; i < n; i++) if (!B.files[i].status) { o1 = B.files[i].o_rgba; break; }
            if (o0 != ~0ull && o1 > o0) {
                cudaEvent_t e = ctx->tex_chunk_ev[chunk % 16];
                UVOL_CUDA(ctx, cudaEventRecord(e, st)); UVOL_CUDA(ctx, cudaStreamWaitEvent(ctx->s4, e, 0));
                UVOL_CUDA(ctx, cudaMemcpyAsync(hO + o0, dO + o0, o1 - o0, cudaMemcpyDeviceToHost, ctx->s4));
            }
        }
        first = false; i0 = i1; chunk++;
    }
    if (hO) { UVOL_CUDA(ctx, cudaEventRecord(ctx->tex_chunk_ev[15], ctx->s4)); UVOL_CUDA(ctx, cudaStreamWaitEvent(st, ctx->tex_chunk_ev[15], 0)); }      // `st` ends after the last result copy
    ctx->span_tex_end = ev - 1;
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_tstate.p, ctx->d_tslices.p, sizeof(TexState) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (ctx->profile && ev < 8) cudaEventRecord(ctx->tex_ev[ev], st);
    ev++;
    B.nev = ev;
    return UVOL_OK;
}

// Replay: the whole resident batch in one go (kernels + result copies), no host sync.
static int ktx2_launch(uvol_ctx *ctx, int memory, cudaStream_t st) {
    TexBatch &B = *ctx->tex; const int n = B.n;
    int ev = 1;
    if (ctx->profile) cudaEventRecord(ctx->tex_ev[0], st);
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_tslices.p, 0, sizeof(TexState) * (size_t)n, st));
    if (ctx->profile && ev < 8) cudaEventRecord(ctx->tex_ev[ev], st);
    ev++;
    B.launches = 0;
    int rc = ktx2_launch_range(ctx, 0, n, st, true, &ev); if (rc) return rc;
    ctx->span_tex_end = ev - 1;
    const size_t st_bytes = align_up(sizeof(TexState) * (size_t)n, 256);
    UVOL_CUDA(ctx, ctx->h_tstate.reserve(st_bytes));
    if (memory == UVOL_MEM_HOST) UVOL_CUDA(ctx, ctx->ph_tout->reserve(B.out + 256));
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_tstate.p, ctx->d_tslices.p, sizeof(TexState) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (memory == UVOL_MEM_HOST) UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->ph_tout->p, ctx->d_out_tex.p, B.out, cudaMemcpyDeviceToHost, st));
    if (ctx->profile && ev < 8) cudaEventRecord(ctx->tex_ev[ev], st);
    ev++;
    B.nev = ev;
    return UVOL_OK;
}

// After the stream has been synchronised: fills the result structs and the statistics.
static int ktx2_finish(uvol_ctx *ctx, int memory, uvol_texture *out, uvol_stats &sx) {
    TexBatch &B = *ctx->tex; const int n = B.n;
    UVOL_CUDA(ctx, cudaGetLastError());
    TexState *hSt = (TexState *)ctx->h_tstate.p; uint8_t *hO = (uint8_t *)ctx->ph_tout->p;
    uint8_t *base = memory == UVOL_MEM_HOST ? hO : (uint8_t *)ctx->d_out_tex.p; uint64_t bytes_out = 0;
    auto out_bytes = [&](const Ktx2File &f) {
        return B.target == UVOL_TEX_ETC1 || B.target == UVOL_TEX_BC1 ? (uint64_t)f.layers * f.bx * f.by * 8 : (B.target == UVOL_TEX_BC7 || B.target == UVOL_TEX_ASTC_4x4 || B.target == UVOL_TEX_ETC2_RGBA || B.target == UVOL_TEX_BC3 ? (uint64_t)f.layers * f.bx * f.by * 16 : (uint64_t)f.layers * f.width * f.height * 4);
    };
    B.mips.assign((size_t)n, uvol_texture_level());          // one entry per internal (single-level) file; a caller's file owns a run of them
    for (int u = 0; u < B.n_user; u++) {
        const uint32_t i0 = B.first_of[u], L = B.nlev[u];
        const Ktx2File &f = B.files[i0]; uvol_texture &t = out[u];
        memset(&t, 0, sizeof t);
        t.status = B.split_status[u];                        // an inconsistent mip chain
        for (uint32_t k = 0; k < L && !t.status; k++) t.status = B.files[i0 + k].status ? B.files[i0 + k].status : hSt[i0 + k].status;      // any failed level fails the texture
        if (t.status) continue;
        t.width = f.width; t.height = f.height; t.layers = f.layers; t.format = (uint32_t)B.target; t.has_alpha = f.has_alpha;
        t.dfd_transfer = f.dfd_transfer; t.dfd_flags = f.dfd_flags;
        t.data = base + f.o_rgba; t.levels = L; t.mips = &B.mips[i0];
        for (uint32_t k = 0; k < L; k++) {
            const Ktx2File &g = B.files[i0 + k]; uvol_texture_level &m = B.mips[i0 + k];
            m.width = g.width; m.height = g.height; m.offset = g.o_rgba - f.o_rgba; m.bytes = out_bytes(g);
            t.bytes = m.offset + m.bytes; bytes_out += m.bytes;
        }
    }
    sx.kernel_launches = B.launches; sx.bytes_in = B.bytes_in; sx.bytes_out = bytes_out; sx.scratch_bytes = B.scratch;
    if (ctx->profile) {
        const int ev = B.nev;
        sx.num_stages = (uint32_t)(ev - 1);
        for (int k = 0; k + 1 < ev && k < 24; k++) cudaEventElapsedTime(&sx.stage_ms[k], ctx->tex_ev[k], ctx->tex_ev[k + 1]);
        float tot = 0; cudaEventElapsedTime(&tot, ctx->tex_ev[1], ctx->tex_ev[ev - 2]); sx.device_ms = tot;
        sx.h2d_ms = sx.stage_ms[0]; sx.d2h_ms = sx.stage_ms[ev - 2];
    }
    return UVOL_OK;
}

static int ktx2_run_replay(uvol_ctx *ctx, int memory, uvol_texture *out) {
    int rc = ktx2_launch(ctx, memory, ctx->s2); if (rc) return rc;
    UVOL_CUDA(ctx, cudaStreamSynchronize(ctx->s2));
    return ktx2_finish(ctx, memory, out, ctx->stats);
}

extern "C" int uvol_transcode_ktx2_batch(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int target_format, int memory, uvol_texture *out) {
    if (!ctx || !out || n < 0 || (n > 0 && (!data || !size)) || n >= (1 << 19)) return UVOL_ERR_ARG;
    if (target_format != UVOL_TEX_RGBA32 && target_format != UVOL_TEX_ETC1 && target_format != UVOL_TEX_BC7 && target_format != UVOL_TEX_ASTC_4x4 && target_format != UVOL_TEX_ETC2_RGBA && target_format != UVOL_TEX_BC1 && target_format != UVOL_TEX_BC3) { ctx->set_error("target formats: UVOL_TEX_RGBA32, UVOL_TEX_ETC1, UVOL_TEX_BC7, UVOL_TEX_ASTC_4x4, UVOL_TEX_ETC2_RGBA, UVOL_TEX_BC1, UVOL_TEX_BC3"); return UVOL_ERR_UNSUPPORTED; }
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->stats, 0, sizeof ctx->stats);
    if (n == 0) { if (ctx->tex) ctx->tex->n = ctx->tex->n_user = 0; return UVOL_OK; }
    const double t0 = now_ms();
    int rc = ktx2_fresh(ctx, data, size, n, target_format, memory, ctx->s2); if (rc) return rc;
    UVOL_CUDA(ctx, cudaStreamSynchronize(ctx->s2));
    rc = ktx2_finish(ctx, memory, out, ctx->stats); if (rc) return rc;
    ctx->stats.host_parse_ms = ctx->tex->parse_ms; ctx->stats.total_ms = now_ms() - t0;
    return UVOL_OK;
}

extern "C" int uvol_replay_ktx2_batch(uvol_ctx *ctx, int memory, uvol_texture *out, int n) {
    if (!ctx || !out || !ctx->tex || ctx->tex->n_user != n || n <= 0) return UVOL_ERR_ARG;
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->stats, 0, sizeof ctx->stats);
    const double t0 = now_ms();
    const int rc = ktx2_run_replay(ctx, memory, out); if (rc) return rc;
    ctx->stats.total_ms = now_ms() - t0;
    return UVOL_OK;
}

// ---- combined V2 step: geometry frames + texture segments decoded concurrently on separate streams,
// like V2Player.fetchBuffers issuing decodeDraco and decodeKTX2 requests to two worker pools at once
// (src/V2/player.ts:272-323).
int uvol_geo_prepare_and_run(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int memory, uvol_geometry *out, bool replay);

extern "C" int uvol_decode_v2_batch(uvol_ctx *ctx, const uint8_t *const *drc, const size_t *drc_size, int n_drc, const uint8_t *const *ktx2, const size_t *ktx2_size,
                                    int n_ktx2, int memory, uvol_geometry *out_geo, uvol_texture *out_tex) {
    if (!ctx || n_drc < 0 || n_ktx2 < 0 || (n_drc && (!drc || !drc_size || !out_geo)) || (n_ktx2 && (!ktx2 || !ktx2_size || !out_tex)) || n_ktx2 >= (1 << 19)) return UVOL_ERR_ARG;
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->stats, 0, sizeof ctx->stats); memset(&ctx->stats_tex, 0, sizeof ctx->stats_tex);
    const double t0 = now_ms();
    int rc = UVOL_OK, rc_tex = UVOL_OK;
    // The texture side (container parse, staging copy of the segments, upload, kernels, result copy on s2) is driven by a helper
    // thread while this thread drives the geometry side, so neither side's host work delays the other's GPU work.
    std::thread tex_thread;
    if (n_ktx2) tex_thread = std::thread([&]() {
        if (cudaSetDevice(ctx->device) != cudaSuccess) { rc_tex = UVOL_ERR_CUDA; return; }
        rc_tex = ktx2_fresh(ctx, ktx2, ktx2_size, n_ktx2, (int)ctx->cfg.texture_target, memory, ctx->s2);
    });
    if (n_drc) rc = uvol_geo_prepare_and_run(ctx, drc, drc_size, n_drc, memory, out_geo, false);
    if (n_ktx2) tex_thread.join();
    if (rc) return rc;
    if (rc_tex) return rc_tex;
    if (n_ktx2) { UVOL_CUDA(ctx, cudaStreamSynchronize(ctx->s2)); rc = ktx2_finish(ctx, memory, out_tex, ctx->stats_tex); if (rc) return rc; ctx->stats_tex.host_parse_ms = ctx->tex->parse_ms; }
    ctx->stats.total_ms = ctx->stats_tex.total_ms = now_ms() - t0;
    return UVOL_OK;
}

// Same, on the batches still resident in HBM (no parse, no input upload).
extern "C" int uvol_replay_v2_batch(uvol_ctx *ctx, int memory, uvol_geometry *out_geo, int n_drc, uvol_texture *out_tex, int n_ktx2) {
    if (!ctx || (n_drc && (!ctx->geo || !out_geo)) || (n_ktx2 && (!ctx->tex || ctx->tex->n_user != n_ktx2 || !out_tex))) return UVOL_ERR_ARG;
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->stats, 0, sizeof ctx->stats); memset(&ctx->stats_tex, 0, sizeof ctx->stats_tex);
    const double t0 = now_ms();
    int rc;
    if (n_ktx2) { rc = ktx2_launch(ctx, memory, ctx->s2); if (rc) return rc; }
    if (n_drc) { rc = uvol_geo_prepare_and_run(ctx, nullptr, nullptr, n_drc, memory, out_geo, true); if (rc) return rc; }
    if (n_ktx2) { UVOL_CUDA(ctx, cudaStreamSynchronize(ctx->s2)); rc = ktx2_finish(ctx, memory, out_tex, ctx->stats_tex); if (rc) return rc; }
    ctx->stats.total_ms = ctx->stats_tex.total_ms = now_ms() - t0;
    return UVOL_OK;
}

extern "C" int uvol_get_stats_kind(const uvol_ctx *c, int kind, uvol_stats *out) { if (!c || !out) return UVOL_ERR_ARG; *out = kind == 1 ? c->stats_tex : c->stats; return UVOL_OK; }
