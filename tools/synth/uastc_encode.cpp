// uastc_encode.cpp -- minimal UASTC LDR 4x4 encoder + KTX2 writer (bench/test INPUT tooling, not product code).
//
// BASELINE.json config C3 needs 2048^2 UASTC KTX2 segments; neither a fixture nor `basisu` exists here
// (SURVEY.md 7.2-1, 8d).  This writes valid UASTC blocks of every mode the caller allows: the mode of a block is picked
// from a hash of its position (so a test can cover all 19 modes), endpoints are the per-subset channel bounds
// quantised to the mode's BISE range, weights the projection of each texel onto that segment, anchor texels are fixed
// up by swapping endpoints.  Quality is not the point -- coverage of the bit layout is.  The transcoding hints are
// filled with hash bits (a real encoder stores ETC1/BC1 hints there; RGBA32 decoding must skip them).
// Container: KTX2, vkFormat 0, DFD colour model 166 (UASTC), supercompression none, layers = frames of the segment.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../universal-volumetric_b200/csrc/uastc_tables.h"

namespace {
typedef std::vector<uint8_t> Bytes;
void put16(Bytes &b, uint32_t v) { b.push_back(v & 255); b.push_back((v >> 8) & 255); }
void put32(Bytes &b, uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((v >> (8 * i)) & 255); }
void put64(Bytes &b, uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((v >> (8 * i)) & 255); }
void pad_to(Bytes &b, size_t a) { while (b.size() % a) b.push_back(0); }

const uint8_t MODE_CODE[19][2] = {{0x01, 4}, {0x35, 6}, {0x1D, 5}, {0x03, 5}, {0x13, 5}, {0x0B, 5}, {0x1B, 5}, {0x07, 5}, {0x17, 5}, {0x0F, 5},
                                  {0x02, 3}, {0x00, 2}, {0x06, 3}, {0x1F, 5}, {0x0D, 5}, {0x05, 7}, {0x15, 6}, {0x25, 6}, {0x09, 4}};
const uint8_t M_COMPS[19]   = {3, 3, 3, 3, 3, 3, 3, 3, 0, 4, 4, 4, 4, 4, 4, 2, 2, 2, 3};
const uint8_t M_SUBSETS[19] = {1, 1, 2, 3, 2, 1, 1, 2, 0, 2, 1, 1, 1, 1, 1, 1, 2, 1, 1};
const uint8_t M_PLANES[19]  = {1, 1, 1, 1, 1, 1, 2, 1, 0, 1, 1, 2, 1, 2, 1, 1, 1, 2, 1};
const uint8_t M_WBITS[19]   = {4, 2, 3, 2, 2, 3, 2, 2, 0, 2, 4, 2, 3, 1, 2, 4, 2, 2, 5};
const uint8_t M_EPROW[19]   = {6, 7, 1, 0, 3, 7, 5, 3, 0, 1, 4, 4, 6, 7, 7, 7, 7, 7, 2};   // row of UASTC_UNQUANT_INIT
const uint8_t M_HINTS[19]   = {15, 15, 15, 15, 15, 15, 15, 15, 0, 23, 17, 17, 17, 23, 23, 23, 23, 23, 15};
const uint8_t ROW_BITS[8] = {2, 4, 5, 3, 4, 5, 6, 8}, ROW_TRITS[8] = {1, 0, 0, 0, 1, 0, 1, 0}, ROW_QUINTS[8] = {0, 0, 0, 1, 0, 1, 0, 0};

struct BitW {
    uint8_t b[16]; uint32_t ofs;
    BitW() : ofs(0) { memset(b, 0, 16); }
    void put(uint32_t v, uint32_t n) { for (uint32_t i = 0; i < n; i++, ofs++) if (ofs < 128 && ((v >> i) & 1u)) b[ofs >> 3] |= (uint8_t)(1u << (ofs & 7)); }
};
uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// nearest raw BISE value of row `row` for an 8-bit target
uint8_t g_nearest[8][256]; bool g_init = false;
void init_tables() {
    if (g_init) return;
    for (int r = 0; r < 8; r++) {
        const int levels = (ROW_TRITS[r] ? 3 : ROW_QUINTS[r] ? 5 : 1) << ROW_BITS[r];
        for (int t = 0; t < 256; t++) { int best = 0, bd = 1 << 30; for (int v = 0; v < levels; v++) { const int d = abs((int)UASTC_UNQUANT_INIT[r][v] - t); if (d < bd) { bd = d; best = v; } } g_nearest[r][t] = (uint8_t)best; }
    }
    g_init = true;
}

void encode_block(const uint8_t px[16][4], uint32_t h, uint32_t mode_mask, uint8_t out[16]) {
    BitW w;
    bool flat = true; for (int i = 1; i < 16; i++) if (memcmp(px[i], px[0], 4)) flat = false;
    int modes[19], nm = 0; for (int m = 0; m < 19; m++) if ((mode_mask >> m) & 1u) if (m != 8) modes[nm++] = m;
    int mode = (flat && ((mode_mask >> 8) & 1u)) || nm == 0 ? 8 : modes[h % (uint32_t)nm];
    w.put(MODE_CODE[mode][0], MODE_CODE[mode][1]);
    if (mode == 8) { for (int c = 0; c < 4; c++) w.put(px[0][c], 8); w.put(mix(h + 1), 21); memcpy(out, w.b, 16); return; }   // colour, then the ETC1 hints
    w.put(mix(h + 2), M_HINTS[mode]);
    const int subsets = M_SUBSETS[mode], planes = M_PLANES[mode], comps = M_COMPS[mode], wbits = M_WBITS[mode], row = M_EPROW[mode], levels = 1 << wbits;
    int part[16] = {0}; uint32_t anchors = 1;
    if (subsets > 1) {
        const uint32_t npat = mode == 3 ? 11 : (mode == 7 ? 19 : 30), pat = mix(h + 3) % npat;
        const uint32_t idx = (mode == 3 ? UASTC_PAT3_BASE : (mode == 7 ? UASTC_PAT7_BASE : 0)) + pat;
        for (int i = 0; i < 16; i++) part[i] = (int)((UASTC_PATTERN_INIT[idx] >> (2 * i)) & 3u);
        anchors = UASTC_ANCHOR_INIT[idx];
        w.put(pat, mode == 3 ? 4 : 5);
    }
    int ccs = -1;
    if (planes == 2) { if (mode == 17) ccs = 3; else { ccs = (int)(mix(h + 4) & 3u); if (comps == 3 && ccs == 3) ccs = 1; w.put((uint32_t)ccs, 2); } }
    // channel values the mode codes: RGB(A) as is, LA = (luma, alpha)
    int v[16][4], nch = comps == 2 ? 2 : comps;
    for (int i = 0; i < 16; i++) {
        if (comps == 2) { v[i][0] = (px[i][0] + px[i][1] + px[i][2] + 1) / 3; v[i][1] = px[i][3]; }
        else for (int c = 0; c < 4; c++) v[i][c] = px[i][c];
    }
    const int ccs_ch = comps == 2 ? (ccs == 3 ? 1 : -1) : ccs;            // index into v[][] of the second-plane channel
    int raw[3][4][2], ql[3][4], qh[3][4]; uint32_t wt[16][2];
    for (int s = 0; s < subsets; s++) {
        for (int c = 0; c < nch; c++) {
            int lo = 255, hi = 0;
            for (int i = 0; i < 16; i++) if (part[i] == s) { lo = std::min(lo, v[i][c]); hi = std::max(hi, v[i][c]); }
            raw[s][c][0] = g_nearest[row][lo]; raw[s][c][1] = g_nearest[row][hi];
            ql[s][c] = UASTC_UNQUANT_INIT[row][raw[s][c][0]]; qh[s][c] = UASTC_UNQUANT_INIT[row][raw[s][c][1]];
        }
    }
    for (int i = 0; i < 16; i++) {
        const int s = part[i];
        for (int p = 0; p < planes; p++) {
            long num = 0, den = 0;
            for (int c = 0; c < nch; c++) {
                if (planes == 2 && ((c == ccs_ch) != (p == 1))) continue;
                const int d = qh[s][c] - ql[s][c]; num += (long)(v[i][c] - ql[s][c]) * d; den += (long)d * d;
            }
            int q = den ? (int)((num * (levels - 1) + den / 2) / den) : 0;
            wt[i][p] = (uint32_t)std::min(levels - 1, std::max(0, q));
        }
    }
    // anchors: the first texel of each subset stores one bit less, so its weight index must be below levels / 2
    for (int s = 0; s < subsets; s++) {
        int a = 0; while (part[a] != s) a++;
        for (int p = 0; p < planes; p++) {
            if (wt[a][p] < (uint32_t)levels / 2) continue;
            for (int i = 0; i < 16; i++) if (part[i] == s) wt[i][p] = (uint32_t)(levels - 1) - wt[i][p];
            for (int c = 0; c < nch; c++) { if (planes == 2 && ((c == ccs_ch) != (p == 1))) continue; std::swap(raw[s][c][0], raw[s][c][1]); }
        }
    }
    // endpoints: bundles of trits / quints first (plain base-3 / base-5 numbers), then the low bits of every value
    const int bits = ROW_BITS[row], tr = ROW_TRITS[row], qu = ROW_QUINTS[row], nvals = comps * 2 * subsets;
    uint32_t vals[18]; int k = 0;
    for (int s = 0; s < subsets; s++) for (int c = 0; c < nch; c++) { vals[k++] = (uint32_t)raw[s][c][0]; vals[k++] = (uint32_t)raw[s][c][1]; }
    if (tr || qu) {
        const int bundle = tr ? 5 : 3, mul = tr ? 3 : 5, ntq = (nvals + bundle - 1) / bundle;
        for (int b = 0; b < ntq; b++) {
            uint32_t acc = 0, m = 1; int cnt = std::min(bundle, nvals - b * bundle);
            for (int j = 0; j < cnt; j++) { acc += (vals[b * bundle + j] >> bits) * m; m *= (uint32_t)mul; }
            int nb = tr ? 8 : 7;
            if (b == ntq - 1) { static const int t[6] = {0, 2, 4, 5, 7, 8}, q[4] = {0, 3, 5, 7}; nb = tr ? t[cnt] : q[cnt]; }
            w.put(acc, (uint32_t)nb);
        }
    }
    for (int i = 0; i < nvals; i++) w.put(vals[i] & ((1u << bits) - 1u), (uint32_t)bits);
    for (int i = 0; i < 16; i++) for (int p = 0; p < planes; p++) w.put(wt[i][p], (uint32_t)(wbits - ((anchors >> i) & 1u)));
    memcpy(out, w.b, 16);
}
}  // namespace



// rgba: layers * h * w * 4.  mode_mask: bit m = mode m may be used (bit 8: solid blocks).  Returns the .ktx2 size.
extern "C" size_t uvsynth_uastc_encode(const uint8_t *rgba, uint32_t w, uint32_t h, uint32_t layers, uint32_t mode_mask, uint32_t seed, int has_alpha, uint8_t **out_buf) {
    *out_buf = nullptr;
    if (!w || !h || !layers) return 0;
    init_tables();
    const uint32_t bx = (w + 3) / 4, by = (h + 3) / 4; const size_t nblk = (size_t)bx * by;
    Bytes level(nblk * 16 * layers);
    for (uint32_t L = 0; L < layers; L++) for (uint32_t yb = 0; yb < by; yb++) for (uint32_t xb = 0; xb < bx; xb++) {
        uint8_t px[16][4];
        for (uint32_t y = 0; y < 4; y++) for (uint32_t x = 0; x < 4; x++) {
            const uint32_t yy = std::min(h - 1, yb * 4 + y), xx = std::min(w - 1, xb * 4 + x);
            memcpy(px[y * 4 + x], rgba + (((size_t)L * h + yy) * w + xx) * 4, 4);
        }
        encode_block(px, mix(seed ^ mix((L * by + yb) * bx + xb)), mode_mask, &level[((size_t)L * nblk + (size_t)yb * bx + xb) * 16]);
    }
    Bytes k; const uint8_t id[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x32, 0x30, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};
    k.insert(k.end(), id, id + 12);
    put32(k, 0); put32(k, 1); put32(k, w); put32(k, h); put32(k, 0); put32(k, layers); put32(k, 1); put32(k, 1); put32(k, 0);
    Bytes kvd;
    { const char key1[] = "KTXanimData"; Bytes v; put32(v, 1); put32(v, 15); put32(v, 0);
      put32(kvd, (uint32_t)(sizeof key1 + v.size())); kvd.insert(kvd.end(), key1, key1 + sizeof key1); kvd.insert(kvd.end(), v.begin(), v.end()); pad_to(kvd, 4);
      const char key2[] = "KTXwriter"; const char val2[] = "uvol-b200 synth (UASTC video)";
      put32(kvd, (uint32_t)(sizeof key2 + sizeof val2)); kvd.insert(kvd.end(), key2, key2 + sizeof key2); kvd.insert(kvd.end(), val2, val2 + sizeof val2); pad_to(kvd, 4); }
    const uint32_t dfdOff = 80 + 24, dfdLen = 44, kvdOff = dfdOff + dfdLen, kvdLen = (uint32_t)kvd.size();
    const uint64_t lvOff = ((uint64_t)kvdOff + kvdLen + 15) / 16 * 16, lvLen = level.size();
    put32(k, dfdOff); put32(k, dfdLen); put32(k, kvdOff); put32(k, kvdLen); put64(k, 0); put64(k, 0);
    put64(k, lvOff); put64(k, lvLen); put64(k, lvLen);
    // DFD: one basic block, colour model 166 (UASTC), one sample of 128 bits; channel id 3 = RGBA, 0 = RGB
    put32(k, 44); put32(k, 0); put16(k, 2); put16(k, 40); k.push_back(166); k.push_back(1); k.push_back(2); k.push_back(0);
    k.push_back(3); k.push_back(3); k.push_back(0); k.push_back(0); k.push_back(16); for (int i = 0; i < 7; i++) k.push_back(0);
    put16(k, 0); k.push_back(127); k.push_back(has_alpha ? 3 : 0); for (int i = 0; i < 4; i++) k.push_back(0); put32(k, 0); put32(k, 0xFFFFFFFFu);
    k.insert(k.end(), kvd.begin(), kvd.end());
    while (k.size() < lvOff) k.push_back(0);
    k.insert(k.end(), level.begin(), level.end());
    *out_buf = (uint8_t *)malloc(k.size() + 16); memcpy(*out_buf, k.data(), k.size());
    return k.size();
}
