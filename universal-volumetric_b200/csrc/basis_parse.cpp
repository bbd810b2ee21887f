// basis_parse.cpp -- host-side structural parse of a .ktx2 container into Ktx2File / Ktx2Slice
// descriptors.  Replaces ktx-parse `read()` + the KTX2File header getters
// (src/lib/ktx-parse.module.js function Pi; src/lib/KTX2Loader.js:299,471-495).  Only the fixed
// header, level index, DFD, key/value block and the BasisLZ image descriptors are touched; the
// codebooks, Huffman tables and slice payloads are decoded on the GPU.  Layout: SURVEY.md B.1/B.2.
#include <string.h>
#include <vector>
#include "uvol_internal.h"

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
