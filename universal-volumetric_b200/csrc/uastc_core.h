// uastc_core.h -- per-block logic of the UASTC LDR 4x4 -> RGBA32 transcode, shared by the sm_100a kernel (uastc_transcode.cu) and
// the host-emulation harness (tests/tools/basis_emu.cpp: logic checks without a GPU; the library has no host transcode).
// Block format: oracle/uastc_oracle.c (the CPU restatement this is checked against bit for bit).
#pragma once
#include <stdint.h>
#include <string.h>
#include "uastc_tables.h"

#if defined(__CUDACC__)
#define UASTC_HD __host__ __device__ __forceinline__
#else
#define UASTC_HD static inline
#endif

// per-mode parameters packed in a word: comps[0:3) subsets[3:5) planes[5:7) wbits[7:10) eprow[10:13) hints[13:18) epbits[18:22) tq[22:24) codelen[24:27)
#define MP(comps, subsets, planes, wbits, eprow, hints, epbits, tq, codelen) \
    ((comps) | ((subsets) << 3) | ((planes) << 5) | ((wbits) << 7) | ((eprow) << 10) | ((hints) << 13) | ((epbits) << 18) | ((tq) << 22) | ((codelen) << 24))
static const uint32_t H_MODE[20] = {
    MP(3, 1, 1, 4, 6, 15, 6, 1, 4), MP(3, 1, 1, 2, 7, 15, 8, 0, 6), MP(3, 2, 1, 3, 1, 15, 4, 0, 5), MP(3, 3, 1, 2, 0, 15, 2, 1, 5),
    MP(3, 2, 1, 2, 3, 15, 3, 2, 5), MP(3, 1, 1, 3, 7, 15, 8, 0, 5), MP(3, 1, 2, 2, 5, 15, 5, 2, 5), MP(3, 2, 1, 2, 3, 15, 3, 2, 5),
    MP(0, 0, 0, 0, 0, 0, 0, 0, 5),  MP(4, 2, 1, 2, 1, 23, 4, 0, 5), MP(4, 1, 1, 4, 4, 17, 4, 1, 3), MP(4, 1, 2, 2, 4, 17, 4, 1, 2),
    MP(4, 1, 1, 3, 6, 17, 6, 1, 3), MP(4, 1, 2, 1, 7, 23, 8, 0, 5), MP(4, 1, 1, 2, 7, 23, 8, 0, 5), MP(2, 1, 1, 4, 7, 23, 8, 0, 7),
    MP(2, 2, 1, 2, 7, 23, 8, 0, 6), MP(2, 1, 2, 2, 7, 23, 8, 0, 6), MP(3, 1, 1, 5, 2, 15, 5, 0, 4), 0};
static const uint8_t H_CODE[20] = {0x01, 0x35, 0x1D, 0x03, 0x13, 0x0B, 0x1B, 0x07, 0x17, 0x0F, 0x02, 0x00, 0x06, 0x1F, 0x0D, 0x05, 0x15, 0x25, 0x09, 0x45};
static const uint8_t H_CODELEN[20] = {4, 6, 5, 5, 5, 5, 5, 5, 5, 5, 3, 2, 3, 5, 5, 7, 6, 6, 4, 7};
static const uint8_t H_WEIGHT[6 * 32] = {      // row = weight bits (row 0 unused)
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 64, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 21, 43, 64, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 9, 18, 27, 37, 46, 55, 64, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 4, 8, 12, 17, 21, 25, 29, 35, 39, 43, 47, 52, 56, 60, 64, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};

struct Bits { uint32_t a, b, c, d; };
UASTC_HD uint32_t uastc_fsr(uint32_t lo, uint32_t hi, uint32_t n) {          // funnel shift right by n in 0..31
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, n);
#else
    return n ? (lo >> n) | (hi << (32u - n)) : lo;
#endif
}
UASTC_HD uint32_t take(Bits &x, uint32_t n) {          // n in 0..31; bits past the block read as zero
    const uint32_t v = x.a & ((1u << n) - 1u);
    x.a = uastc_fsr(x.a, x.b, n); x.b = uastc_fsr(x.b, x.c, n); x.c = uastc_fsr(x.c, x.d, n); x.d >>= n;
    return v;
}

struct UastcShared {
    uint32_t mode[20]; uint32_t pattern[60]; uint16_t anchor[60]; uint8_t mode_of[128]; uint8_t weight[6 * 32]; uint8_t unquant[8 * 256]; uint8_t pad[8];
};
static_assert(sizeof(UastcShared) % 16 == 0, "staged with one bulk copy (16-byte granules)");

// One block -> four pixel rows of packed RGBA.  false: the transcoder rejects the block.
UASTC_HD bool uastc_block(const UastcShared &T, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t rows[4][4]) {
    Bits x{w0, w1, w2, w3};
    const uint32_t mode = T.mode_of[x.a & 127u];
    if (mode >= 19u) return false;
    const uint32_t mp = T.mode[mode];
    take(x, mp >> 24);
    if (mode == 8u) {
        const uint32_t c = x.a;          // R, G, B, A bytes follow the mode code
#pragma unroll
        for (int y = 0; y < 4; y++) { rows[y][0] = c; rows[y][1] = c; rows[y][2] = c; rows[y][3] = c; }
        return true;
    }
    const uint32_t comps = mp & 7u, subsets = (mp >> 3) & 3u, planes = (mp >> 5) & 3u, wbits = (mp >> 7) & 7u, eprow = (mp >> 10) & 7u,
                   epbits = (mp >> 18) & 15u, tq = (mp >> 22) & 3u;
    take(x, (mp >> 13) & 31u);                                             // transcoding hints: not needed for RGBA32
    uint32_t pattern = 0, anchors = 1;
    if (subsets > 1u) {
        const uint32_t three = mode == 3u, pat = take(x, three ? 4u : 5u);
        const uint32_t limit = three ? 11u : (mode == 7u ? 19u : 30u);
        if (pat >= limit) return false;
        const uint32_t idx = pat + (three ? UASTC_PAT3_BASE : (mode == 7u ? UASTC_PAT7_BASE : 0));
        pattern = T.pattern[idx]; anchors = T.anchor[idx];
    }
    uint32_t ccs = 4;                                                      // channel on the second weight plane (4: none)
    if (planes == 2u) ccs = mode == 17u ? 3u : take(x, 2);
    // ---- endpoints: trit / quint bundles first (plain base-3 / base-5 numbers), then the low bits of each value
    const uint32_t nvals = comps * 2u * subsets;
    uint32_t tqpack = 0;
    if (tq) {
        const uint32_t bundle = tq == 1u ? 5u : 3u, full = tq == 1u ? 8u : 7u, ntq = (nvals + bundle - 1u) / bundle, rem = nvals - (ntq - 1u) * bundle;
        const uint32_t last = tq == 1u ? ((0x875420u >> (4u * rem)) & 15u) : ((0x7530u >> (4u * rem)) & 15u);    // bits of a partial bundle
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) if (i < ntq) tqpack |= take(x, i == ntq - 1u ? last : full) << (8u * i);
    }
    const uint32_t mul = tq == 1u ? 3u : 5u, bundle = tq == 1u ? 5u : 3u;
    uint32_t lo[3], hi[3], accum = 0, left = 0;
    const uint32_t init = comps == 3u ? 0xff000000u : 0u;
    const uint8_t *unq = T.unquant + eprow * 256u;
#pragma unroll
    for (uint32_t s = 0; s < 3; s++) {
        lo[s] = init; hi[s] = init;
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
#pragma unroll
            for (uint32_t e = 0; e < 2; e++) {
                if (s < subsets && c < comps) {
                    uint32_t v = take(x, epbits);
                    if (tq) {
                        if (!left) { accum = tqpack & 255u; tqpack >>= 8; left = bundle; }
                        const uint32_t q = tq == 1u ? (accum * 171u) >> 9 : (accum * 205u) >> 10;
                        v |= (accum - q * mul) << epbits; accum = q; left--;
                    }
                    const uint32_t u = unq[v & 255u];
                    const uint32_t put = comps == 2u ? (c == 0 ? u * 0x010101u : u << 24) : u << (8u * c);
                    if (e == 0) lo[s] |= put; else hi[s] |= put;
                }
            }
        }
    }
    // ---- weights + interpolation, texel by texel (the first texel of every subset stores one bit less)
    const uint8_t *wtab = T.weight + wbits * 32u;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t s = (pattern >> (2 * i)) & 3u;
        const uint32_t l = s == 0 ? lo[0] : (s == 1 ? lo[1] : lo[2]), h = s == 0 ? hi[0] : (s == 1 ? hi[1] : hi[2]);
        const uint32_t nb = wbits - ((anchors >> i) & 1u);
        const uint32_t w0 = wtab[take(x, nb)];
        uint32_t w1 = w0;
        if (planes == 2u) w1 = wtab[take(x, nb)];
        // ASTC interpolation ((l*257*(64-w) + h*257*w + 32) >> 6) >> 8 == (t + ((t + 32) >> 8)) >> 6 with t = l*(64-w) + h*w <= 16320,
        // evaluated for two channels at a time in 16-bit halves (R|B and G|A); the second-plane channel is patched in afterwards
        const uint32_t lrb = l & 0x00ff00ffu, lga = (l >> 8) & 0x00ff00ffu, hrb = h & 0x00ff00ffu, hga = (h >> 8) & 0x00ff00ffu;
        uint32_t trb = lrb * (64u - w0) + hrb * w0, tga = lga * (64u - w0) + hga * w0;
        trb = ((trb + (((trb + 0x00200020u) >> 8) & 0x00ff00ffu)) >> 6) & 0x00ff00ffu;
        tga = ((tga + (((tga + 0x00200020u) >> 8) & 0x00ff00ffu)) >> 6) & 0x00ff00ffu;
        uint32_t px = trb | (tga << 8);
        if (planes == 2u) {
            const uint32_t lc = (l >> (8u * ccs)) & 255u, hc = (h >> (8u * ccs)) & 255u, t = lc * (64u - w1) + hc * w1;
            px = (px & ~(255u << (8u * ccs))) | (((t + ((t + 32u) >> 8)) >> 6) << (8u * ccs));
        }
        rows[i >> 2][i & 3] = px;
    }
    return true;
}

// Host: the table image (per-mode words, mode-prefix lookup, partition patterns, anchors, weight and endpoint unquantisation).
static inline void uastc_fill_tables(UastcShared &h) {
    memset(&h, 0, sizeof h);
    memcpy(h.mode, H_MODE, sizeof h.mode); memcpy(h.pattern, UASTC_PATTERN_INIT, sizeof h.pattern); memcpy(h.anchor, UASTC_ANCHOR_INIT, sizeof h.anchor);
    memcpy(h.weight, H_WEIGHT, sizeof h.weight); memcpy(h.unquant, UASTC_UNQUANT_INIT, sizeof h.unquant);
    for (uint32_t v = 0; v < 128; v++) { uint32_t m = 19; for (uint32_t k = 0; k < 20; k++) if ((v & ((1u << H_CODELEN[k]) - 1u)) == H_CODE[k]) m = k; h.mode_of[v] = (uint8_t)m; }
}
