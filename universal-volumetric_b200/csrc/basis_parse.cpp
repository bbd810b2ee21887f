// basis_parse.cpp -- host-side structural parse of a .ktx2 container into Ktx2File / Ktx2Slice
// descriptors.  Replaces ktx-parse `read()` + the KTX2File header getters
// (src/lib/ktx-parse.module.js function Pi; src/lib/KTX2Loader.js:299,471-495).  Only the fixed
// header, level index, DFD, key/value block and the BasisLZ image descriptors are touched; the
// codebooks, Huffman tables and slice payloads are decoded on the GPU.  Layout: SURVEY.md B.1/B.2.
#include <string.h>
#include <algorithm>
#include <vector>
#include "uvol_internal.h"
#include "../../include/uvol_b200.h"

namespace {
uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
const uint8_t KTX2_ID[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x32, 0x30, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};
}

int uvol_ktx2_parse(const uint8_t *b, size_t len, uint32_t file_index, Ktx2File &f, std::vector<Ktx2Slice> &slices) {
    if (len < 104 || memcmp(b, KTX2_ID, 12)) return UVOL_ERR_CORRUPT;
    const uint32_t vk = rd32(b + 12), w = rd32(b + 20), h = rd32(b + 24), depth = rd32(b + 28), layers = rd32(b + 32), faces = rd32(b + 36);
    uint32_t levels = rd32(b + 40); const uint32_t sc = rd32(b + 44);
    const uint32_t dfdOff = rd32(b + 48), dfdLen = rd32(b + 52), kvdOff = rd32(b + 56), kvdLen = rd32(b + 60);
    const uint64_t sgdOff = rd64(b + 64), sgdLen = rd64(b + 72);
    if (vk != 0 || depth != 0 || faces != 1 || w == 0 || h == 0 || w > 16384 || h > 16384) return UVOL_ERR_UNSUPPORTED;
    if (levels == 0) levels = 1;
    if (levels != 1) return UVOL_ERR_UNSUPPORTED;      // UVOL textures carry no mips (scripts/Encoder.py:290)
    const uint32_t nl = layers ? layers : 1;
    // every (offset, length) pair comes from the file: compare without forming offset + length (a 64-bit sum wraps)
    auto inside = [len](uint64_t off, uint64_t n) { return off <= len && n <= len - off; };
    if (!inside(dfdOff, dfdLen) || !inside(kvdOff, kvdLen) || !inside(sgdOff, sgdLen) || dfdLen < 44) return UVOL_ERR_CORRUPT;
    const uint64_t lvOff = rd64(b + 80), lvLen = rd64(b + 88);
    if (!inside(lvOff, lvLen) || lvOff > 0xffffffffull) return UVOL_ERR_TRUNCATED;
    const int color_model = b[dfdOff + 12];
    f.dfd_transfer = b[dfdOff + 14]; f.dfd_flags = b[dfdOff + 15];
    const int nsamples = (int)((rd16(b + dfdOff + 10) - 24) / 16);
    const int chan0 = b[dfdOff + 28 + 3] & 0xF;
    f.width = w; f.height = h; f.layers = nl; f.bx = (w + 3) / 4; f.by = (h + 3) / 4;
    f.is_uastc = color_model == 166; f.is_video = 0;
    for (uint64_t p = kvdOff, end = (uint64_t)kvdOff + kvdLen; end - p >= 4;) {          // 64-bit cursor: keyAndValueByteLength is attacker-controlled
        const uint64_t kl = rd32(b + p); p += 4;
        if (kl > end - p) break;
        if (kl >= 11 && !memcmp(b + p, "KTXanimData", 11)) f.is_video = 1;
        const uint64_t adv = (kl + 3) & ~3ull;                                           // always forward, never past the block
        if (adv == 0 || adv > end - p) break;
        p += adv;
    }
    f.level_off = (uint32_t)lvOff; f.first_slice = (uint32_t)slices.size();
    const uint64_t nblk = (uint64_t)f.bx * f.by;
    f.zstd = 0;
    if (f.is_uastc) {
        if (sc != 0 && sc != 2) return UVOL_ERR_UNSUPPORTED;      // none or Zstandard (KTX2Loader.js:46-47); zlib (3) is not used by basisu
        f.has_alpha = chan0 == 3;
        if (sc == 2) {          // the level is one Zstandard frame: inflated by the host into the staging blob (csrc/zstd_inflate.cpp)
            const uint64_t ulen = rd64(b + 96);
            // the inflated level is exactly the block payload: a header asking for more than that is rejected before anything is reserved
            if (ulen != (uint64_t)nl * nblk * 16 || ulen >= (1ull << 31) || lvLen >= (1ull << 31) || lvLen == 0) return ulen >= (1ull << 31) ? UVOL_ERR_UNSUPPORTED : UVOL_ERR_TRUNCATED;
            f.zstd = 1; f.z_src_off = (uint32_t)lvOff; f.z_src_len = (uint32_t)lvLen; f.z_len = (uint32_t)ulen; f.level_off = 0;
            f.endpoint_count = f.selector_count = 0;
            return UVOL_OK;
        }
        if (lvLen < (uint64_t)nl * nblk * 16) return UVOL_ERR_TRUNCATED;
        f.endpoint_count = f.selector_count = 0;
        return UVOL_OK;
    }
    if (color_model != 163 || sc != 1) return UVOL_ERR_UNSUPPORTED;
    f.has_alpha = nsamples == 2;
    if (sgdLen < 20 + 20ull * nl) return UVOL_ERR_CORRUPT;
    const uint8_t *g = b + sgdOff;
    const uint32_t ec = rd16(g), scnt = rd16(g + 2), eb = rd32(g + 4), sb = rd32(g + 8), tb = rd32(g + 12);
    if (20 + 20ull * nl + eb + sb + tb > sgdLen || ec == 0 || scnt == 0 || sgdOff + sgdLen > 0xffffffffull) return UVOL_ERR_CORRUPT;
    f.endpoint_count = ec; f.selector_count = scnt;
    f.ep_off = (uint32_t)(sgdOff + 20 + 20ull * nl); f.ep_len = eb;
    f.sel_off = f.ep_off + eb; f.sel_len = sb; f.tab_off = f.sel_off + sb; f.tab_len = tb;
    for (int plane = 0; plane < (f.has_alpha ? 2 : 1); plane++) {
        for (uint32_t L = 0; L < nl; L++) {
            const uint8_t *d = g + 20 + 20 * L;
            const uint32_t off = rd32(d + 4 + 8 * plane), ln = rd32(d + 8 + 8 * plane);
            if (off > lvLen || ln > lvLen - off || ln == 0 || lvOff + off + ln > len) return UVOL_ERR_TRUNCATED;
            Ktx2Slice s; memset(&s, 0, sizeof s);
            s.file = file_index; s.layer = L; s.data_off = (uint32_t)(lvOff + off); s.data_len = ln; s.is_alpha = (uint32_t)plane;
            slices.push_back(s);
        }
    }
    return UVOL_OK;
}

// Mip levels (levelCount > 1; the level / layer / face loop of src/lib/KTX2Loader.js:514-573).  The transcode machinery works on
// single-level files, so a file with a mip chain is taken apart HERE, on the host, into one synthetic single-level KTX2 per level
// (level 0 = the largest first): header with that level's dimensions and levelCount 1, a one-entry level index, the DFD and the
// key/value block as they are, for BasisLZ the global data with the image descriptors of that level only (codebooks and tables
// copied), then the level's payload (still Zstandard-compressed if the file's levels are).  Image descriptors are ordered level-major
// (level 0 first), then by layer.  Returns the number of levels written to `out`, 0 when the file has a single level or is not a
// KTX2 file this function understands (the caller passes it on unchanged and the parser reports on it), or a negative status when
// the mip chain itself is inconsistent.  Cube faces (faceCount 6) are not taken apart: such files stay UNSUPPORTED.
int uvol_ktx2_split_levels(const uint8_t *b, size_t len, std::vector<std::vector<uint8_t>> &out) {
    if (!b || len < 104 || memcmp(b, KTX2_ID, 12)) return 0;
    const uint32_t w = rd32(b + 20), h = rd32(b + 24), layers = rd32(b + 32), faces = rd32(b + 36), levels = rd32(b + 40), sc = rd32(b + 44);
    if (levels <= 1 || faces != 1) return 0;
    if (levels > 15 || w == 0 || h == 0 || w > 16384 || h > 16384 || ((w >> (levels - 1)) == 0 && (h >> (levels - 1)) == 0)) return UVOL_ERR_CORRUPT;
    auto inside = [len](uint64_t off, uint64_t n) { return off <= len && n <= len - off; };
    if (len < 80 + 24ull * levels) return UVOL_ERR_TRUNCATED;
    const uint32_t dfdOff = rd32(b + 48), dfdLen = rd32(b + 52), kvdOff = rd32(b + 56), kvdLen = rd32(b + 60);
    const uint64_t sgdOff = rd64(b + 64), sgdLen = rd64(b + 72);
    if (!inside(dfdOff, dfdLen) || !inside(kvdOff, kvdLen) || !inside(sgdOff, sgdLen)) return UVOL_ERR_CORRUPT;
    const uint32_t nl = layers ? layers : 1;
    uint64_t cb_bytes = 0;                                   // BasisLZ: codebooks + tables (+ extended data) behind the image descriptors
    if (sc == 1) {
        if (sgdLen < 20 + 20ull * nl * levels) return UVOL_ERR_CORRUPT;
        cb_bytes = sgdLen - 20 - 20ull * nl * levels;
    }
    for (uint32_t k = 0; k < levels; k++) {
        const uint64_t lvOff = rd64(b + 80 + 24ull * k), lvLen = rd64(b + 88 + 24ull * k), lvU = rd64(b + 96 + 24ull * k);
        if (!inside(lvOff, lvLen) || lvLen >= (1ull << 31)) return UVOL_ERR_TRUNCATED;
        const uint64_t o_dfd = 104, o_kvd = (o_dfd + dfdLen + 3) & ~3ull, o_sgd = (o_kvd + kvdLen + 7) & ~7ull;
        const uint64_t sgd2 = sc == 1 ? 20 + 20ull * nl + cb_bytes : 0, o_lv = (o_sgd + sgd2 + 15) & ~15ull;
        std::vector<uint8_t> f((size_t)(o_lv + lvLen), 0);
        memcpy(f.data(), b, 80);
        const uint32_t lw = std::max(1u, w >> k), lh = std::max(1u, h >> k), one = 1;
        auto w32 = [&](size_t at, uint32_t v) { memcpy(f.data() + at, &v, 4); };
        auto w64 = [&](size_t at, uint64_t v) { memcpy(f.data() + at, &v, 8); };
        w32(20, lw); w32(24, lh); w32(40, one);
        w32(48, (uint32_t)o_dfd); w32(56, kvdLen ? (uint32_t)o_kvd : 0); w64(64, sgd2 ? o_sgd : 0); w64(72, sgd2);
        w64(80, o_lv); w64(88, lvLen); w64(96, lvU);
        memcpy(f.data() + o_dfd, b + dfdOff, dfdLen);
        if (kvdLen) memcpy(f.data() + o_kvd, b + kvdOff, kvdLen);
        if (sc == 1) {
            const uint8_t *g = b + sgdOff;
            memcpy(f.data() + o_sgd, g, 20);
            memcpy(f.data() + o_sgd + 20, g + 20 + 20ull * nl * k, 20ull * nl);
            memcpy(f.data() + o_sgd + 20 + 20ull * nl, g + 20 + 20ull * nl * levels, (size_t)cb_bytes);
        }
        memcpy(f.data() + o_lv, b + lvOff, (size_t)lvLen);
        out.push_back(std::move(f));
    }
    return (int)levels;
}

// KTX2File's header getters (src/lib/KTX2Loader.js:471-495) for a caller that picks the target format before transcoding.
extern "C" int uvol_ktx2_probe(const uint8_t *data, size_t size, uvol_ktx2_info *out) {
    if (!data || !out) return UVOL_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (size < 104 || memcmp(data, KTX2_ID, 12)) return UVOL_ERR_CORRUPT;
    out->levels = std::max(1u, rd32(data + 40)); out->faces = rd32(data + 36); out->supercompression = rd32(data + 44);
    std::vector<std::vector<uint8_t>> lv; std::vector<Ktx2Slice> slices; Ktx2File f; memset(&f, 0, sizeof f);
    const uint8_t *b = data; size_t n = size;
    if (out->levels > 1) {          // a mip chain is described by its base level
        const int L = uvol_ktx2_split_levels(data, size, lv);
        if (L < 0) return L;
        if (L > 0) { b = lv[0].data(); n = lv[0].size(); }
    }
    const int rc = uvol_ktx2_parse(b, n, 0, f, slices);
    if (rc) return rc;
    out->width = f.width; out->height = f.height; out->layers = f.layers; out->is_uastc = f.is_uastc; out->has_alpha = f.has_alpha; out->is_video = f.is_video;
    out->dfd_transfer = f.dfd_transfer; out->dfd_flags = f.dfd_flags;
    return UVOL_OK;
}
