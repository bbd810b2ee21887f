// astc_core.h -- per-block logic of the ASTC 4x4 target format (UVOL_TEX_ASTC_4x4): UASTC LDR 4x4 -> ASTC LDR 4x4, shared by the
// sm_100a kernel (uastc_transcode.cu) and the host-emulation harness (tests/tools/basis_emu.cpp).
//
// Replaces transcodeImage(..., ASTC_4x4, ...) -- the first entry of the reference's FORMAT_OPTIONS, chosen for UASTC sources when
// the GPU has WEBGL_compressed_texture_astc (src/lib/KTX2Loader.js:592-600; UASTC only: priorityETC1S is Infinity, so ETC1S sources
// with this target report UVOL_STATUS_UNSUPPORTED per item, as the reference would fall through to its next option).
// UASTC is a subset of ASTC: every mode is one ASTC configuration (block mode, partition count and seed, colour endpoint mode, endpoint
// range), so the repack is LOSSLESS -- the ASTC block decodes to exactly the texels of the RGBA32 path.  It is checked bit for bit
// through an independent ASTC decoder (oracle/astc_decode.c, which derives the endpoint range from the bit budget the way hardware
// does) against the oracle's UASTC decode (tests/test_astc.py).  What changes between the two containers:
//   * endpoint values keep their bits; their trits / quints are re-bundled from plain base-3 / base-5 numbers into ASTC's trit /
//     quint blocks, interleaved with the value bits (integer sequence encoding);
//   * ASTC's RGB / RGBA direct endpoint modes switch to "blue contraction" when the second endpoint's R+G+B sum is below the first's;
//     UASTC never means that, so such a subset has its endpoint pairs swapped and its weights complemented (same texels);
//   * weights are stored in full (UASTC drops the top bit of each subset's first texel), bit-reversed from the top of the block, the
//     second-plane channel selector right below them;
//   * solid-colour blocks become void-extent blocks.
#pragma once
#include "uastc_core.h"

struct AstcShared { uint16_t seed[60]; uint16_t block_mode[20]; uint8_t trit[243]; uint8_t quint[125]; };
static_assert(sizeof(AstcShared) % 16 == 0, "staged with one bulk copy (16-byte granules)");

// ASTC block mode (11 bits) of each UASTC mode: 4x4 weight grid (bits 2-3 = 00, A = 2 at bits 5-6, B = 0 at bits 7-8), weight range
// in R (bits 4, 0, 1) and the precision bit 9, bit 10 = two weight planes.  1 bit: 0x41, 2: 0x42, 3: 0x53, 4: 0x242, 5: 0x253.
static inline void astc_fill_tables(AstcShared &h) {
    memset(&h, 0, sizeof h);
    memcpy(h.seed, UASTC_ASTC_SEED_INIT, sizeof h.seed); memcpy(h.trit, ASTC_TRIT_ENC_INIT, sizeof h.trit); memcpy(h.quint, ASTC_QUINT_ENC_INIT, sizeof h.quint);
    static const uint16_t by_wbits[6] = {0, 0x41, 0x42, 0x53, 0x242, 0x253};
    for (uint32_t m = 0; m < 19; m++) {
        const uint32_t mp = H_MODE[m], planes = (mp >> 5) & 3u, wbits = (mp >> 7) & 7u;
        h.block_mode[m] = m == 8 ? 0 : (uint16_t)(by_wbits[wbits] | (planes == 2u ? 0x400u : 0u));
    }
}

// 128-bit LSB-first writer (two 64-bit halves)
struct AstcBits { unsigned long long lo, hi; uint32_t pos; };
UASTC_HD void astc_put(AstcBits &b, uint32_t v, uint32_t n) {          // n in 0..16
    if (b.pos < 64u) { b.lo |= (unsigned long long)v << b.pos; if (b.pos + n > 64u) b.hi |= (unsigned long long)v >> (64u - b.pos); }
    else b.hi |= (unsigned long long)v << (b.pos - 64u);
    b.pos += n;
}
UASTC_HD unsigned long long astc_rev64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __brevll(v);
#else
    v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4) & 0x0f0f0f0f0f0f0f0full) | ((v & 0x0f0f0f0f0f0f0f0full) << 4);
    return __builtin_bswap64(v);
#endif
}

UASTC_HD int astc_ctz(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
// inserts a zero bit at position p (0..126) of the 128-bit value {lo, hi}: bits p.. move up by one
UASTC_HD void astc_ins0(unsigned long long &lo, unsigned long long &hi, uint32_t p) {
    if (p < 64u) { const unsigned long long m = (1ull << p) - 1ull; hi = (hi << 1) | (lo >> 63); lo = (lo & m) | ((lo & ~m) << 1); }
    else { const unsigned long long m = (1ull << (p - 64u)) - 1ull; hi = (hi & m) | ((hi & ~m) << 1); }
}

// One UASTC block -> one ASTC block (four little-endian words).  false: the block is rejected (like uastc_block).
UASTC_HD bool uastc_to_astc(const UastcShared &T, const AstcShared &A, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, uint32_t out[4]) {
    Bits x{q0, q1, q2, q3};
    const uint32_t mode = T.mode_of[x.a & 127u];
    if (mode >= 19u) return false;
    const uint32_t mp = T.mode[mode];
    take(x, mp >> 24);
    if (mode == 8u) {          // void extent, LDR: 0x1FC, reserved bits set, no extent (all ones), four UNORM16 components
        const uint32_t c = x.a;
        out[0] = 0xFFFFFDFCu; out[1] = 0xFFFFFFFFu;
        out[2] = ((c & 255u) * 257u) | ((((c >> 8) & 255u) * 257u) << 16);
        out[3] = (((c >> 16) & 255u) * 257u) | (((c >> 24) * 257u) << 16);
        return true;
    }
    const uint32_t comps = mp & 7u, subsets = (mp >> 3) & 3u, planes = (mp >> 5) & 3u, wbits = (mp >> 7) & 7u, eprow = (mp >> 10) & 7u,
                   epbits = (mp >> 18) & 15u, tq = (mp >> 22) & 3u;
    take(x, (mp >> 13) & 31u);                                             // transcoding hints (for other targets)
    uint32_t pattern = 0, anchors = 1, seed = 0;
    if (subsets > 1u) {
        const uint32_t three = mode == 3u, pat = take(x, three ? 4u : 5u);
        const uint32_t limit = three ? 11u : (mode == 7u ? 19u : 30u);
        if (pat >= limit) return false;
        const uint32_t idx = pat + (three ? UASTC_PAT3_BASE : (mode == 7u ? UASTC_PAT7_BASE : 0));
        pattern = T.pattern[idx]; anchors = T.anchor[idx]; seed = A.seed[idx];
    }
    uint32_t ccs = 0;
    if (planes == 2u) ccs = mode == 17u ? 3u : take(x, 2);
    // ---- endpoint values: bits | digit << 8 per slot (subset, component, end), and the R+G+B sums that decide blue contraction
    const uint32_t nvals = comps * 2u * subsets;
    uint32_t tqpack = 0;
    if (tq) {
        const uint32_t bundle = tq == 1u ? 5u : 3u, full = tq == 1u ? 8u : 7u, ntq = (nvals + bundle - 1u) / bundle, rem = nvals - (ntq - 1u) * bundle;
        const uint32_t last = tq == 1u ? ((0x875420u >> (4u * rem)) & 15u) : ((0x7530u >> (4u * rem)) & 15u);
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) if (i < ntq) tqpack |= take(x, i == ntq - 1u ? last : full) << (8u * i);
    }
    const uint32_t mul = tq == 1u ? 3u : 5u, bundle = tq == 1u ? 5u : 3u;
    uint32_t slot[3][4][2], sum[3][2], accum = 0, left = 0;
    const uint8_t *unq = T.unquant + eprow * 256u;
#pragma unroll
    for (uint32_t s = 0; s < 3; s++) {
        sum[s][0] = sum[s][1] = 0;
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
#pragma unroll
            for (uint32_t e = 0; e < 2; e++) {
                slot[s][c][e] = 0;
                if (s < subsets && c < comps) {
                    const uint32_t v = take(x, epbits); uint32_t d = 0;
                    if (tq) {
                        if (!left) { accum = tqpack & 255u; tqpack >>= 8; left = bundle; }
                        const uint32_t q = tq == 1u ? (accum * 171u) >> 9 : (accum * 205u) >> 10;
                        d = accum - q * mul; accum = q; left--;
                    }
                    slot[s][c][e] = v | (d << 8);
                    if (c < 3u) sum[s][e] += unq[(v | (d << epbits)) & 255u];
                }
            }
        }
    }
    // (a bundle value beyond its digits' range is ignored like the RGBA32 path does: the digits themselves are always in range)
    bool swap[3];
#pragma unroll
    for (uint32_t s = 0; s < 3; s++) swap[s] = comps >= 3u && sum[s][1] < sum[s][0];
    // ---- header + endpoint values (integer sequence encoding: value bits interleaved with pieces of the trit / quint block)
    AstcBits b{0, 0, 0};
    astc_put(b, A.block_mode[mode], 11); astc_put(b, subsets - 1u, 2);
    const uint32_t cem = comps == 3u ? 8u : (comps == 4u ? 12u : 4u);
    if (subsets > 1u) { astc_put(b, seed, 10); astc_put(b, cem << 2, 6); } else astc_put(b, cem, 4);
    uint32_t blocks = 0;                                                   // the trit / quint block of each bundle, 8 bits apiece, in emission order
    if (tq) {
        uint32_t num = 0, pw = 1, k = 0, nb = 0;
#pragma unroll
        for (uint32_t s = 0; s < 3; s++)
#pragma unroll
            for (uint32_t c = 0; c < 4; c++)
#pragma unroll
                for (uint32_t e = 0; e < 2; e++) if (s < subsets && c < comps) {
                    const uint32_t sl = swap[s] ? slot[s][c][e ^ 1u] : slot[s][c][e];
                    num += (sl >> 8) * pw; pw *= mul;
                    if (++k == bundle) { blocks |= (uint32_t)(tq == 1u ? A.trit[num] : A.quint[num]) << (8u * nb); nb++; num = 0; pw = 1; k = 0; }
                }
        if (k) blocks |= (uint32_t)(tq == 1u ? A.trit[num] : A.quint[num]) << (8u * nb);
    }
    {
        const uint32_t widths = tq == 1u ? 0x12122u : 0x223u;             // bits of the block that follow value k of a bundle
        uint32_t cur = blocks & 255u, k = 0;
        blocks >>= 8;
#pragma unroll
        for (uint32_t s = 0; s < 3; s++)
#pragma unroll
            for (uint32_t c = 0; c < 4; c++)
#pragma unroll
                for (uint32_t e = 0; e < 2; e++) if (s < subsets && c < comps) {
                    const uint32_t sl = swap[s] ? slot[s][c][e ^ 1u] : slot[s][c][e];
                    uint32_t v = sl & 255u, nbits = epbits;
                    if (tq) {                       // the piece of the trit / quint block that follows this value goes out with it
                        const uint32_t n = (widths >> (4u * k)) & 15u;
                        v |= (cur & ((1u << n) - 1u)) << epbits; nbits += n; cur >>= n;
                        if (++k == bundle) { k = 0; cur = blocks & 255u; blocks >>= 8; }
                    }
                    astc_put(b, v, nbits);
                }
    }
    // ---- weights: the stored stream (what is left of the block, right-aligned) is turned into ASTC's as a whole -- a zero bit
    // inserted above each subset's first weight (which UASTC stores one bit short), the fields of swapped subsets complemented by
    // one XOR, the second-plane selector appended -- and then mirrored onto the top of the block
    unsigned long long wl = (unsigned long long)x.a | ((unsigned long long)x.b << 32), wh = (unsigned long long)x.c | ((unsigned long long)x.d << 32);
    const uint32_t stride = wbits * planes, wtotal = 16u * stride;
    {
        uint32_t am = anchors;
#pragma unroll
        for (int t = 0; t < 3; t++) if (am) {
            const uint32_t i = (uint32_t)astc_ctz(am); am &= am - 1u;
            astc_ins0(wl, wh, i * stride + wbits - 1u);
            if (planes == 2u) astc_ins0(wl, wh, i * stride + 2u * wbits - 1u);
        }
    }
    if (swap[0] | swap[1] | swap[2]) {
        unsigned long long ml = 0, mh = 0;
        if (subsets == 1u) { ml = ~0ull; mh = ~0ull; }
        else {
            const unsigned long long ones = (1ull << stride) - 1ull;          // stride <= 3 with more than one subset: the whole field lies in the low word
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t sb = (pattern >> (2 * i)) & 3u;
                if (sb == 0 ? swap[0] : (sb == 1 ? swap[1] : swap[2])) ml |= ones << ((uint32_t)i * stride);
            }
        }
        wl ^= ml; wh ^= mh;
    }
    // keep exactly the weight field, then the selector (bit-swapped: the stream is mirrored, the selector is not)
    if (wtotal < 64u) { wl &= (1ull << wtotal) - 1ull; wh = 0; } else if (wtotal < 128u) wh &= (1ull << (wtotal - 64u)) - 1ull;
    if (planes == 2u) { const unsigned long long sel = ((ccs & 1u) << 1) | (ccs >> 1); if (wtotal < 64u) wl |= sel << wtotal; else wh |= sel << (wtotal - 64u); }
    const unsigned long long lo = b.lo | astc_rev64(wh), hi = b.hi | astc_rev64(wl);
    out[0] = (uint32_t)lo; out[1] = (uint32_t)(lo >> 32); out[2] = (uint32_t)hi; out[3] = (uint32_t)(hi >> 32);
    return true;
}
