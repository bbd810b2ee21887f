// basis_emu.cpp -- HOST EMULATION of the texture pipeline's per-unit logic (basis_core.h) for
// logic checks without a GPU.  Test tool only; never part of libuvol_b200.so, never a fallback.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../universal-volumetric_b200/csrc/uvol_internal.h"
#include "../../universal-volumetric_b200/csrc/basis_core.h"
#include "../../universal-volumetric_b200/csrc/uastc_core.h"
#include "../../universal-volumetric_b200/csrc/bc7_core.h"
#include "../../universal-volumetric_b200/csrc/astc_core.h"

extern "C" int uvol_zstd_inflate(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len);

int uvol_ktx2_parse(const uint8_t *b, size_t len, uint32_t file_index, Ktx2File &f, std::vector<Ktx2Slice> &slices);
int uvol_ktx2_split_levels(const uint8_t *b, size_t len, std::vector<std::vector<uint8_t>> &out);

// The product's mip-chain splitter (basis_parse.cpp): returns its result; level k's synthetic single-level file is copied to
// files[k] (malloc'ed, sizes[k] bytes), at most `cap` of them.
extern "C" int basis_emu_split_levels(const uint8_t *data, size_t len, uint8_t **files, size_t *sizes, int cap) {
    std::vector<std::vector<uint8_t>> out;
    const int rc = uvol_ktx2_split_levels(data, len, out);
    for (int k = 0; k < rc && k < cap; k++) { files[k] = (uint8_t *)malloc(out[k].size()); memcpy(files[k], out[k].data(), out[k].size()); sizes[k] = out[k].size(); }
    return rc;
}

static int basis_emu_decode_impl(const uint8_t *data, size_t len, uint8_t **rgba, uint32_t *w, uint32_t *h, uint32_t *layers, int etc1);
extern "C" int basis_emu_decode(const uint8_t *data, size_t len, uint8_t **rgba, uint32_t *w, uint32_t *h, uint32_t *layers) { return basis_emu_decode_impl(data, len, rgba, w, h, layers, 0); }
// target ETC1: *rgba receives layers * blocks * 8 bytes (opaque ETC1S files only)
extern "C" int basis_emu_decode_etc1(const uint8_t *data, size_t len, uint8_t **blocks, uint32_t *w, uint32_t *h, uint32_t *layers) { return basis_emu_decode_impl(data, len, blocks, w, h, layers, 1); }
// target BC7: *blocks receives layers * blocks * 16 bytes (ETC1S with or without alpha, UASTC)
extern "C" int basis_emu_decode_bc7(const uint8_t *data, size_t len, uint8_t **blocks, uint32_t *w, uint32_t *h, uint32_t *layers) { return basis_emu_decode_impl(data, len, blocks, w, h, layers, 2); }
// targets BC1 / BC3: *blocks receives layers * blocks * 8 / 16 bytes (ETC1S files)
extern "C" int basis_emu_decode_bc1(const uint8_t *data, size_t len, uint8_t **blocks, uint32_t *w, uint32_t *h, uint32_t *layers) { return basis_emu_decode_impl(data, len, blocks, w, h, layers, 5); }
extern "C" int basis_emu_decode_bc3(const uint8_t *data, size_t len, uint8_t **blocks, uint32_t *w, uint32_t *h, uint32_t *layers) { return basis_emu_decode_impl(data, len, blocks, w, h, layers, 6); }
// target ETC2 RGBA: *blocks receives layers * blocks * 16 bytes (ETC1S files, with or without alpha)
extern "C" int basis_emu_decode_etc2a(const uint8_t *data, size_t len, uint8_t **blocks, uint32_t *w, uint32_t *h, uint32_t *layers) { return basis_emu_decode_impl(data, len, blocks, w, h, layers, 4); }
// target ASTC 4x4: *blocks receives layers * blocks * 16 bytes (UASTC sources only, like the reference's ASTC option)
extern "C" int basis_emu_decode_astc(const uint8_t *data, size_t len, uint8_t **blocks, uint32_t *w, uint32_t *h, uint32_t *layers) { return basis_emu_decode_impl(data, len, blocks, w, h, layers, 3); }
static int basis_emu_decode_impl(const uint8_t *data, size_t len, uint8_t **rgba, uint32_t *w, uint32_t *h, uint32_t *layers, int etc1) {
    std::vector<uint8_t> padded(len + 64, 0); memcpy(padded.data(), data, len);   // the launcher pads the blob the same way
    const uint8_t *file = padded.data();
    Ktx2File f; memset(&f, 0, sizeof f); std::vector<Ktx2Slice> slices;
    int rc = uvol_ktx2_parse(file, len, 0, f, slices); if (rc) return rc;
    if (f.is_uastc && (etc1 == 1 || etc1 == 4 || etc1 == 5 || etc1 == 6)) return UVOL_ERR_UNSUPPORTED;
    if (!f.is_uastc && etc1 == 3) return UVOL_ERR_UNSUPPORTED;
    if (f.is_uastc) {          // the kernel's per-block function (uastc_core.h) over every block, Zstd levels inflated by the product's decoder
        const uint32_t nblk = f.bx * f.by;
        std::vector<uint8_t> inflated; const uint8_t *level = file + f.level_off;
        if (f.zstd) {
            inflated.resize(f.z_len); size_t got = 0;
            rc = uvol_zstd_inflate(file + f.z_src_off, f.z_src_len, inflated.data(), inflated.size(), &got);
            if (rc || got != f.z_len) return rc ? rc : UVOL_ERR_CORRUPT;
            level = inflated.data();
        }
        UastcShared T; uastc_fill_tables(T);
        if (etc1 == 2) {          // the BC7 kernel's per-block function (bc7_core.h)
            Bc7Shared B7; bc7_fill_tables(B7);
            uint8_t *out = (uint8_t *)malloc((size_t)f.layers * nblk * 16 + 16);
            for (uint32_t L = 0; L < f.layers; L++) for (uint32_t bi = 0; bi < nblk; bi++) {
                uint32_t w[4], o[4]; memcpy(w, level + ((size_t)L * nblk + bi) * 16, 16);
                if (!uastc_to_bc7(T, B7, w[0], w[1], w[2], w[3], o)) { free(out); return UVOL_ERR_CORRUPT; }
                memcpy(out + ((size_t)L * nblk + bi) * 16, o, 16);
            }
            *rgba = out; *w = f.width; *h = f.height; *layers = f.layers;
            return UVOL_OK;
        }
        if (etc1 == 3) {          // the ASTC kernel's per-block function (astc_core.h)
            AstcShared AS; astc_fill_tables(AS);
            uint8_t *out = (uint8_t *)malloc((size_t)f.layers * nblk * 16 + 16);
            for (uint32_t L = 0; L < f.layers; L++) for (uint32_t bi = 0; bi < nblk; bi++) {
                uint32_t w[4], o[4]; memcpy(w, level + ((size_t)L * nblk + bi) * 16, 16);
                if (!uastc_to_astc(T, AS, w[0], w[1], w[2], w[3], o)) { free(out); return UVOL_ERR_CORRUPT; }
                memcpy(out + ((size_t)L * nblk + bi) * 16, o, 16);
            }
            *rgba = out; *w = f.width; *h = f.height; *layers = f.layers;
            return UVOL_OK;
        }
        uint8_t *out = (uint8_t *)malloc((size_t)f.layers * f.width * f.height * 4 + 16);
        for (uint32_t L = 0; L < f.layers; L++) for (uint32_t bi = 0; bi < nblk; bi++) {
            uint32_t w[4]; memcpy(w, level + ((size_t)L * nblk + bi) * 16, 16);
            uint32_t rows[4][4];
            if (!uastc_block(T, w[0], w[1], w[2], w[3], rows)) { free(out); return UVOL_ERR_CORRUPT; }
            const uint32_t xb = bi % f.bx, yb = bi / f.bx;
            for (uint32_t y = 0; y < 4 && yb * 4 + y < f.height; y++) for (uint32_t x = 0; x < 4 && xb * 4 + x < f.width; x++)
                memcpy(out + (((size_t)L * f.height + yb * 4 + y) * f.width + xb * 4 + x) * 4, &rows[y][x], 4);
        }
        *rgba = out; *w = f.width; *h = f.height; *layers = f.layers;
        return UVOL_OK;
    }
    const uint32_t pool_cap = f.endpoint_count + f.selector_count + 8192 + 1024, nblk = f.bx * f.by;
    std::vector<uint32_t> eps(f.endpoint_count), sels(f.selector_count); std::vector<HuffTable> tabs(10);
    std::vector<uint16_t> pool(pool_cap + 16384 + 64);
    BasisGlobalsMem m{eps.data(), sels.data(), tabs.data(), pool.data(), (uint8_t *)(pool.data() + pool_cap), pool_cap};
    uint32_t hs = 0;
    rc = basis_build_globals(f, file, m, &hs); if (rc) return rc;
    const size_t ns = slices.size();
    std::vector<std::vector<uint8_t>> pred(ns); std::vector<std::vector<uint16_t>> delta(ns), sel(ns), ep(ns);
    std::vector<uint8_t> rowp(4096); std::vector<uint16_t> hist(1024);
    for (size_t k = 0; k < ns; k++) {
        pred[k].resize(nblk); delta[k].resize(nblk); sel[k].resize(nblk); ep[k].resize(nblk);
        BitRd b; br_init(b, file + slices[k].data_off, slices[k].data_len);
        SliceTables T{&tabs[0], &tabs[1], &tabs[2], &tabs[3], pool.data()};
        rc = etc1s_slice_symbols(b, T, f.bx, f.by, f.selector_count, hs, (int)f.is_video, rowp.data(), hist.data(), pred[k].data(), delta[k].data(), sel[k].data());
        if (rc) return rc;
        if ((b.consumed + 7) / 8 != slices[k].data_len) return UVOL_ERR_TRUNCATED;
    }
    // serial resolve (the kernel does the same with a warp segmented scan)
    for (uint32_t plane = 0; plane < (f.has_alpha ? 2u : 1u); plane++) for (uint32_t L = 0; L < f.layers; L++) {
        const size_t k = plane * f.layers + L; uint32_t prev = 0;
        for (uint32_t y = 0; y < f.by; y++) for (uint32_t x = 0; x < f.bx; x++) {
            const uint32_t bi = y * f.bx + x, p = pred[k][bi]; uint32_t e;
            if (p == 0) e = prev; else if (p == 1) e = ep[k][bi - f.bx];
            else if (p == 2) { if (f.is_video) { if (!L) return UVOL_ERR_CORRUPT; e = ep[k - 1][bi]; sel[k][bi] = sel[k - 1][bi]; } else e = ep[k][bi - f.bx - 1]; }
            else { e = delta[k][bi] + prev; if (e >= f.endpoint_count) e -= f.endpoint_count; }
            ep[k][bi] = (uint16_t)e; prev = e;
        }
    }
    *w = f.width; *h = f.height; *layers = f.layers;
    if (etc1 == 2) {            // the BC7 kernel's per-block function (bc7_core.h etc1s_to_bc7)
        Bc7Shared B7; bc7_fill_tables(B7);
        *rgba = (uint8_t *)malloc((size_t)f.layers * nblk * 16 + 16);
        for (uint32_t L = 0; L < f.layers; L++) for (uint32_t bi = 0; bi < nblk; bi++) {
            uint32_t o[4];
            if (f.has_alpha) etc1s_to_bc7(B7, eps[ep[L][bi]], sels[sel[L][bi]], true, eps[ep[f.layers + L][bi]], sels[sel[f.layers + L][bi]], o);
            else etc1s_to_bc7(B7, eps[ep[L][bi]], sels[sel[L][bi]], false, 0, 0, o);
            memcpy(*rgba + ((size_t)L * nblk + bi) * 16, o, 16);
        }
        return 0;
    }
    if (etc1 == 5 || etc1 == 6) {          // the BC1 / BC3 kernel's per-block functions (basis_core.h etc1s_to_bc1 + etc1s_alpha_to_bc4)
        const size_t bs = etc1 == 5 ? 8 : 16;
        *rgba = (uint8_t *)malloc((size_t)f.layers * nblk * bs + 16);
        for (uint32_t L = 0; L < f.layers; L++) for (uint32_t bi = 0; bi < nblk; bi++) {
            const Bc1Words c = etc1s_to_bc1(eps[ep[L][bi]], sels[sel[L][bi]]);
            uint8_t *o = *rgba + ((size_t)L * nblk + bi) * bs;
            if (etc1 == 6) {
                const Bc1Words a = f.has_alpha ? etc1s_alpha_to_bc4(eps[ep[f.layers + L][bi]], sels[sel[f.layers + L][bi]]) : bc4_opaque();
                memcpy(o, &a.x, 4); memcpy(o + 4, &a.y, 4); o += 8;
            }
            memcpy(o, &c.x, 4); memcpy(o + 4, &c.y, 4);
        }
        return 0;
    }
    if (etc1 == 4) {            // the ETC2 RGBA kernel's per-block functions (basis_core.h etc1s_alpha_to_eac + etc1s_to_etc1)
        *rgba = (uint8_t *)malloc((size_t)f.layers * nblk * 16 + 16);
        for (uint32_t L = 0; L < f.layers; L++) for (uint32_t bi = 0; bi < nblk; bi++) {
            const Etc1Words e = etc1s_to_etc1(eps[ep[L][bi]], sels[sel[L][bi]]);
            const EacWords a = f.has_alpha ? etc1s_alpha_to_eac(eps[ep[f.layers + L][bi]], sels[sel[f.layers + L][bi]], ETC1S_EAC_MAP_INIT) : eac_opaque();
            uint8_t *o = *rgba + ((size_t)L * nblk + bi) * 16;
            memcpy(o, &a.x, 4); memcpy(o + 4, &a.y, 4); memcpy(o + 8, &e.x, 4); memcpy(o + 12, &e.y, 4);
        }
        return 0;
    }
    if (etc1) {                 // the kernel's repack function (basis_core.h etc1s_to_etc1) over every block
        if (f.has_alpha) return UVOL_ERR_UNSUPPORTED;
        *rgba = (uint8_t *)malloc((size_t)f.layers * nblk * 8 + 8);
        for (uint32_t L = 0; L < f.layers; L++) for (uint32_t bi = 0; bi < nblk; bi++) {
            const Etc1Words e = etc1s_to_etc1(eps[ep[L][bi]], sels[sel[L][bi]]);
            memcpy(*rgba + ((size_t)L * nblk + bi) * 8, &e.x, 4); memcpy(*rgba + ((size_t)L * nblk + bi) * 8 + 4, &e.y, 4);
        }
        return 0;
    }
    *rgba = (uint8_t *)malloc((size_t)f.layers * f.width * f.height * 4);
    for (uint32_t L = 0; L < f.layers; L++) for (uint32_t bi = 0; bi < nblk; bi++) {
        uint32_t rows[4][4]; etc1s_block_rows(eps[ep[L][bi]], sels[sel[L][bi]], rows);
        if (f.has_alpha) {          // as k_etc1s_blocks: alpha = G of the alpha slice's block colour
            const uint32_t aep = eps[ep[f.layers + L][bi]], ase = sels[sel[f.layers + L][bi]];
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) { const uint32_t q = (ase >> (8 * y + 2 * x)) & 3u; rows[y][x] = (rows[y][x] & 0x00ffffffu) | (((etc1s_color(aep, (int)q) >> 8) & 255u) << 24); }
        }
        const uint32_t xb = bi % f.bx, yb = bi / f.bx;
        for (uint32_t y = 0; y < 4 && yb * 4 + y < f.height; y++) for (uint32_t x = 0; x < 4 && xb * 4 + x < f.width; x++)
            memcpy(*rgba + ((size_t)L * f.width * f.height + (size_t)(yb * 4 + y) * f.width + xb * 4 + x) * 4, &rows[y][x], 4);
    }
    return 0;
}
