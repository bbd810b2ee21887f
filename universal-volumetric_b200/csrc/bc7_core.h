// bc7_core.h -- per-block logic of the BC7 target format (UVOL_TEX_BC7): UASTC LDR 4x4 -> BC7 and ETC1S -> BC7 mode 5, shared by
// the sm_100a kernels and the host-emulation harness (tests/tools/basis_emu.cpp).
//
// Replaces transcodeImage(..., BC7_M5, ...) -- the format the reference's FORMAT_OPTIONS pick for ETC1S *and* UASTC sources on a
// desktop GPU with EXT_texture_compression_bptc (src/lib/KTX2Loader.js:602-604, selection :659-689).  The reference's arithmetic for
// this lives in the absent basis_transcoder WASM, whose ETC1S path goes through large precomputed tables; bit-matching it is not
// possible here, so this is an own transcoder built on UASTC's design property -- every UASTC mode has a BC7 mode
// with the same subset shapes and weight grid -- and it is validated by an independent BC7 decoder (oracle/bc7_decode.c): exact
// for solid blocks, bounded error elsewhere (tests/test_bc7.py states the bounds).
//
//   UASTC mode                                   -> BC7 mode
//   8  solid colour                              -> 5  (endpoint pair that reproduces every 8-bit value exactly at index 1; alpha 8 bit)
//   6, 11, 13, 17  one subset, two weight planes -> 5  (rotation = the second-plane channel, its weights = the scalar indices)
//   0, 1, 5, 10, 12, 14, 15, 18  one subset      -> 6  (7777 + p per endpoint, 4-bit indices; opaque sources force p = 1 so alpha stays 255)
//   2  two subsets, 3-bit weights                -> 1  (666 + shared p, 3-bit indices)
//   4  two subsets, 2-bit weights (RGB)          -> 3  (777 + p, 2-bit indices)
//   9, 16  two subsets with alpha                -> 7  (5555 + p, 2-bit indices)
//   3  three subsets                             -> 2  (555, 2-bit indices)
//   7  two subsets on a three-subset BC7 shape   -> 2  (two BC7 subsets share one UASTC subset's endpoints)
// Weight indices map one to one when the grids agree (2- and 3-bit: identical weights) and to the nearest BC7 weight otherwise.
#pragma once
#include "uastc_core.h"
#include "basis_core.h"

// Endpoint quantisation is table driven (one shared-memory byte per value): q5 / q7 = nearest 5- / 7-bit code under MSB replication;
// q7p[p] / q6p[p] = nearest 6- / 5-bit code whose value with the p-bit p appended (7 / 6 bits in all) is nearest, bit 7 of the p = 0
// row = the parity of the unconstrained nearest code (the vote for the shared p-bit).  7 bits + p is the value itself.
struct Bc7Shared { uint32_t info[60]; uint32_t pat3[20]; uint8_t wmap[3][6][32]; uint8_t solid5[256][2]; uint8_t q5[256], q7[256], q7p[2][256], q6p[2][256]; };
static_assert(sizeof(Bc7Shared) % 16 == 0, "staged with one bulk copy (16-byte granules)");
static inline uint32_t bc7_expand_host(uint32_t x, uint32_t bits) { x <<= (8u - bits); return (x | (x >> bits)) & 255u; }
static inline uint32_t bc7_nearest_host(uint32_t e, uint32_t bits, int parity) {
    uint32_t best = 0, beste = 0xffffu;
    for (uint32_t x = 0; x < (1u << bits); x++) {
        if (parity >= 0 && (int)(x & 1u) != parity) continue;
        const uint32_t ex = bc7_expand_host(x, bits), er = ex > e ? ex - e : e - ex;
        if (er < beste) { beste = er; best = x; }
    }
    return best;
}
static inline void bc7_fill_tables(Bc7Shared &h) {
    memset(&h, 0, sizeof h);
    memcpy(h.info, UASTC_BC7_INFO_INIT, sizeof UASTC_BC7_INFO_INIT); memcpy(h.pat3, UASTC_BC7_PAT3_INIT, sizeof UASTC_BC7_PAT3_INIT);
    memcpy(h.wmap, UASTC_BC7_WMAP_INIT, sizeof h.wmap); memcpy(h.solid5, BC7_SOLID5_INIT, sizeof h.solid5);
    for (uint32_t e = 0; e < 256; e++) {
        h.q5[e] = (uint8_t)bc7_nearest_host(e, 5, -1); h.q7[e] = (uint8_t)bc7_nearest_host(e, 7, -1);
        for (int p = 0; p < 2; p++) { h.q7p[p][e] = (uint8_t)(bc7_nearest_host(e, 7, p) >> 1); h.q6p[p][e] = (uint8_t)(bc7_nearest_host(e, 6, p) >> 1); }
        h.q7p[0][e] |= (uint8_t)((bc7_nearest_host(e, 7, -1) & 1u) << 7); h.q6p[0][e] |= (uint8_t)((bc7_nearest_host(e, 6, -1) & 1u) << 7);
    }
}

// ---- 128-bit LSB-first writer held in two 64-bit registers
struct Bc7Bits { unsigned long long lo, hi; uint32_t pos; };
UASTC_HD void b7_put(Bc7Bits &b, uint32_t v, uint32_t n) {          // n in 1..16
    if (b.pos < 64u) { b.lo |= (unsigned long long)v << b.pos; if (b.pos + n > 64u) b.hi |= (unsigned long long)v >> (64u - b.pos); }
    else b.hi |= (unsigned long long)v << (b.pos - 64u);
    b.pos += n;
}
UASTC_HD uint32_t b7_expand(uint32_t x, uint32_t bits) { x <<= (8u - bits); return (x | (x >> bits)) & 255u; }
UASTC_HD uint32_t b7_absdiff(uint32_t a, uint32_t b) { return a > b ? a - b : b - a; }
UASTC_HD uint32_t b7_interp(uint32_t a, uint32_t b, uint32_t w) { return ((64u - w) * a + w * b + 32u) >> 6; }

// ---- the logical content of a UASTC block
struct UastcLogical {
    uint32_t mode, comps, subsets, planes, wbits, ccs, pat_idx, pattern, solid;
    uint32_t lo[3], hi[3];          // packed RGBA8 endpoints per UASTC subset
    uint8_t w0[16], w1[16];         // weight indices of plane 0 / plane 1
};
UASTC_HD bool uastc_unpack(const UastcShared &T, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, UastcLogical &L) {
    Bits x{q0, q1, q2, q3};
    const uint32_t mode = T.mode_of[x.a & 127u];
    if (mode >= 19u) return false;
    const uint32_t mp = T.mode[mode];
    take(x, mp >> 24);
    L.mode = mode; L.solid = 0; L.pattern = 0; L.pat_idx = 0; L.ccs = 4;
    if (mode == 8u) { L.solid = x.a; L.comps = 4; L.subsets = 1; L.planes = 1; L.wbits = 0; return true; }
    const uint32_t comps = mp & 7u, subsets = (mp >> 3) & 3u, planes = (mp >> 5) & 3u, wbits = (mp >> 7) & 7u, eprow = (mp >> 10) & 7u,
                   epbits = (mp >> 18) & 15u, tq = (mp >> 22) & 3u;
    L.comps = comps; L.subsets = subsets; L.planes = planes; L.wbits = wbits;
    take(x, (mp >> 13) & 31u);                                             // basisu's own transcoding hints: this transcoder derives everything from the block
    uint32_t anchors = 1;
    if (subsets > 1u) {
        const uint32_t three = mode == 3u, pat = take(x, three ? 4u : 5u);
        const uint32_t limit = three ? 11u : (mode == 7u ? 19u : 30u);
        if (pat >= limit) return false;
        const uint32_t idx = pat + (three ? UASTC_PAT3_BASE : (mode == 7u ? UASTC_PAT7_BASE : 0));
        L.pattern = T.pattern[idx]; anchors = T.anchor[idx]; L.pat_idx = idx;
    }
    if (planes == 2u) L.ccs = mode == 17u ? 3u : take(x, 2);
    const uint32_t nvals = comps * 2u * subsets;
    uint32_t tqpack = 0;
    if (tq) {
        const uint32_t bundle = tq == 1u ? 5u : 3u, full = tq == 1u ? 8u : 7u, ntq = (nvals + bundle - 1u) / bundle, rem = nvals - (ntq - 1u) * bundle;
        const uint32_t last = tq == 1u ? ((0x875420u >> (4u * rem)) & 15u) : ((0x7530u >> (4u * rem)) & 15u);
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) if (i < ntq) tqpack |= take(x, i == ntq - 1u ? last : full) << (8u * i);
    }
    const uint32_t mul = tq == 1u ? 3u : 5u, bundle = tq == 1u ? 5u : 3u;
    uint32_t accum = 0, left = 0;
    const uint32_t init = comps == 3u ? 0xff000000u : 0u;
    const uint8_t *unq = T.unquant + eprow * 256u;
#pragma unroll
    for (uint32_t s = 0; s < 3; s++) {
        L.lo[s] = init; L.hi[s] = init;
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
                    if (e == 0) L.lo[s] |= put; else L.hi[s] |= put;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t nb = wbits - ((anchors >> i) & 1u);
        L.w0[i] = (uint8_t)take(x, nb);
        L.w1[i] = planes == 2u ? (uint8_t)take(x, nb) : L.w0[i];
    }
    return true;
}

// ---- BC7 mode 5 (one subset; RGB 7 bits + scalar 8 bits, two 2-bit index sets, rotation)
// c0 / c1: 7-bit colour endpoints (packed, one per byte), a0 / a1: 8-bit scalar endpoints, ci / ai: 2 bits per texel
UASTC_HD void bc7_pack_mode5(uint32_t rot, uint32_t c0, uint32_t c1, uint32_t a0, uint32_t a1, uint32_t ci, uint32_t ai, uint32_t out[4]) {
    if (ci & 2u) { const uint32_t t = c0; c0 = c1; c1 = t; ci = ~ci; }          // anchor (texel 0) keeps its high bit clear
    if (ai & 2u) { const uint32_t t = a0; a0 = a1; a1 = t; ai = ~ai; }
    Bc7Bits b{0, 0, 0};
    b7_put(b, 1u << 5, 6); b7_put(b, rot, 2);
#pragma unroll
    for (uint32_t c = 0; c < 3; c++) { b7_put(b, (c0 >> (8u * c)) & 127u, 7); b7_put(b, (c1 >> (8u * c)) & 127u, 7); }
    b7_put(b, a0 & 255u, 8); b7_put(b, a1 & 255u, 8);
    b7_put(b, ci & 1u, 1); b7_put(b, (ci >> 2) & 0x3fffu, 14); b7_put(b, (ci >> 16) & 0xffffu, 16);
    b7_put(b, ai & 1u, 1); b7_put(b, (ai >> 2) & 0x3fffu, 14); b7_put(b, (ai >> 16) & 0xffffu, 16);
    out[0] = (uint32_t)b.lo; out[1] = (uint32_t)(b.lo >> 32); out[2] = (uint32_t)b.hi; out[3] = (uint32_t)(b.hi >> 32);
}

// ---- the partitioned / single-subset modes 1, 2, 3, 6, 7
// e0[s] / e1[s]: 8-bit RGBA endpoints of BC7 subset s (unused subsets: anything); sub: BC7 subset per texel (2 bits each);
// idx: BC7 index per texel, 4 bits each
UASTC_HD void bc7_pack_general(const Bc7Shared &B, uint32_t m, uint32_t part, uint32_t anc1, uint32_t anc2, const uint32_t e0[3], const uint32_t e1[3], uint32_t sub,
                               unsigned long long idx, bool opaque, uint32_t out[4]) {
    const uint32_t ns = m == 6u ? 1u : (m == 2u ? 3u : 2u), cb = m == 1u ? 6u : (m == 2u ? 5u : (m == 7u ? 5u : 7u)), ab = m == 6u ? 7u : (m == 7u ? 5u : 0u),
                   ib = m == 1u ? 3u : (m == 6u ? 4u : 2u), ptype = m == 2u ? 0u : (m == 1u ? 2u : 1u);      // p-bits: none / per endpoint / shared per subset
    const uint8_t *tp0 = m == 1u ? B.q7p[0] : B.q6p[0], *tp1 = m == 1u ? B.q7p[1] : B.q6p[1];
    const uint32_t nch = ab ? 4u : 3u;
    uint32_t q[3][2], pb[3][2];                                    // quantised endpoints, one channel per byte
#pragma unroll
    for (uint32_t s = 0; s < 3; s++) {
        const uint32_t E[2] = {e0[s], e1[s]};
        uint32_t votes[2];
#pragma unroll
        for (uint32_t e = 0; e < 2; e++) {                         // the parity most channels' nearest codes have
            if (cb == 7u) votes[e] = (uint32_t)__builtin_popcount(E[e] & (ab ? 0x01010101u : 0x00010101u));
            else { votes[e] = 0; for (uint32_t c = 0; c < 4; c++) if (c < nch) votes[e] += tp0[(E[e] >> (8u * c)) & 255u] >> 7; }
        }
#pragma unroll
        for (uint32_t e = 0; e < 2; e++) {
            uint32_t p = 0;
            if (ptype == 1u) { p = votes[e] * 2u >= nch ? 1u : 0u; if (ab && (opaque || (E[e] >> 24) == 255u)) p = 1u; }      // an alpha of 255 must stay 255: only the all-ones code reaches it
            else if (ptype == 2u) p = votes[0] + votes[1] >= 3u ? 1u : 0u;
            uint32_t r = 0;
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) {
                const uint32_t v = (E[e] >> (8u * c)) & 255u; uint32_t qq;
                if (ptype == 0u) qq = B.q5[v];
                else if (cb == 7u) { uint32_t x = v; if ((x & 1u) != p) x = v == 255u ? 254u : v + 1u; qq = x >> 1; }      // 7 bits + p: the value itself, moved by one when the low bit disagrees
                else qq = (p ? tp1[v] : tp0[v]) & 127u;
                r |= qq << (8u * c);
            }
            q[s][e] = r; pb[s][e] = p;
        }
    }
    // anchors: the index of each subset's anchor texel is stored without its high bit; a subset whose anchor has it set is flipped
    // (endpoints swapped, indices complemented)
    const uint32_t msb = 1u << (ib - 1u), maxi = (1u << ib) - 1u;
    const uint32_t f0 = (uint32_t)(idx & 15u) & msb, f1 = (uint32_t)((idx >> (4u * anc1)) & 15u) & msb, f2 = (uint32_t)((idx >> (4u * anc2)) & 15u) & msb;
    const bool flip[3] = {f0 != 0u, ns > 1u && f1 != 0u, ns > 2u && f2 != 0u};
#pragma unroll
    for (uint32_t s = 0; s < 3; s++) if (flip[s]) { const uint32_t t = q[s][0]; q[s][0] = q[s][1]; q[s][1] = t; const uint32_t u = pb[s][0]; pb[s][0] = pb[s][1]; pb[s][1] = u; }
#pragma unroll
    for (int i = 0; i < 16; i++) { const uint32_t sb = (sub >> (2 * i)) & 3u; if (sb == 0u ? flip[0] : (sb == 1u ? flip[1] : flip[2])) idx ^= (unsigned long long)maxi << (4 * i); }
    Bc7Bits b{0, 0, 0};
    b7_put(b, 1u << m, m + 1u);
    if (ns > 1u) b7_put(b, part, 6);
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        if (c == 3u && !ab) continue;
        const uint32_t nbits = c == 3u ? ab : cb;
#pragma unroll
        for (uint32_t s = 0; s < 3; s++) if (s < ns) { b7_put(b, (q[s][0] >> (8u * c)) & 255u, nbits); b7_put(b, (q[s][1] >> (8u * c)) & 255u, nbits); }
    }
    if (ptype == 1u) { for (uint32_t s = 0; s < 3; s++) if (s < ns) { b7_put(b, pb[s][0], 1); b7_put(b, pb[s][1], 1); } }
    else if (ptype == 2u) { for (uint32_t s = 0; s < 3; s++) if (s < ns) b7_put(b, pb[s][0], 1); }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const bool is_anchor = (uint32_t)i == 0u || (ns > 1u && (uint32_t)i == anc1) || (ns > 2u && (uint32_t)i == anc2);
        b7_put(b, (uint32_t)(idx >> (4 * i)) & maxi, is_anchor ? ib - 1u : ib);
    }
    out[0] = (uint32_t)b.lo; out[1] = (uint32_t)(b.lo >> 32); out[2] = (uint32_t)b.hi; out[3] = (uint32_t)(b.hi >> 32);
}

// One UASTC block -> one BC7 block (four little-endian words).  false: the block is rejected (like uastc_block).
UASTC_HD bool uastc_to_bc7(const UastcShared &T, const Bc7Shared &B, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, uint32_t out[4]) {
    UastcLogical L;
    if (!uastc_unpack(T, q0, q1, q2, q3, L)) return false;
    if (L.mode == 8u) {
        const uint32_t c = L.solid; uint32_t c0 = 0, c1 = 0;
#pragma unroll
        for (uint32_t k = 0; k < 3; k++) { const uint32_t v = (c >> (8u * k)) & 255u; c0 |= (uint32_t)B.solid5[v][0] << (8u * k); c1 |= (uint32_t)B.solid5[v][1] << (8u * k); }
        bc7_pack_mode5(0, c0, c1, c >> 24, c >> 24, 0x55555555u, 0u, out);
        return true;
    }
    if (L.planes == 2u) {          // dual plane -> mode 5, rotation = the channel on the second plane
        const uint32_t ccs = L.ccs, rot = ccs == 3u ? 0u : ccs + 1u;
        uint32_t l = L.lo[0], h = L.hi[0];
        const uint32_t s0 = (l >> (8u * ccs)) & 255u, s1 = (h >> (8u * ccs)) & 255u;      // scalar endpoints: 8 bits, exact
        if (rot) {                 // the alpha endpoints move into the colour slot of the rotated channel
            l = (l & ~(255u << (8u * ccs))) | ((l >> 24) << (8u * ccs)); h = (h & ~(255u << (8u * ccs))) | ((h >> 24) << (8u * ccs));
        }
        uint32_t c0 = 0, c1 = 0;
#pragma unroll
        for (uint32_t k = 0; k < 3; k++) { c0 |= (uint32_t)B.q7[(l >> (8u * k)) & 255u] << (8u * k); c1 |= (uint32_t)B.q7[(h >> (8u * k)) & 255u] << (8u * k); }
        uint32_t ci = 0, ai = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) { ci |= (uint32_t)B.wmap[0][L.wbits][L.w0[i]] << (2 * i); ai |= (uint32_t)B.wmap[0][L.wbits][L.w1[i]] << (2 * i); }
        bc7_pack_mode5(rot, c0, c1, s0, s1, ci, ai, out);
        return true;
    }
    const bool opaque = L.comps == 3u;
    uint32_t m, part = 0, anc1 = 0, anc2 = 0, sub = 0, e0[3], e1[3];
    if (L.subsets == 1u) { m = 6; e0[0] = L.lo[0]; e1[0] = L.hi[0]; e0[1] = e0[2] = e1[1] = e1[2] = 0; }
    else {
        const uint32_t inf = B.info[L.pat_idx];
        part = inf & 63u; anc1 = (inf >> 9) & 15u; anc2 = (inf >> 13) & 15u;
        const uint32_t src0 = (inf >> 17) & 3u, src1 = (inf >> 19) & 3u, src2 = (inf >> 21) & 3u;
        e0[0] = src0 == 0 ? L.lo[0] : (src0 == 1 ? L.lo[1] : L.lo[2]); e1[0] = src0 == 0 ? L.hi[0] : (src0 == 1 ? L.hi[1] : L.hi[2]);
        e0[1] = src1 == 0 ? L.lo[0] : (src1 == 1 ? L.lo[1] : L.lo[2]); e1[1] = src1 == 0 ? L.hi[0] : (src1 == 1 ? L.hi[1] : L.hi[2]);
        e0[2] = src2 == 0 ? L.lo[0] : (src2 == 1 ? L.lo[1] : L.lo[2]); e1[2] = src2 == 0 ? L.hi[0] : (src2 == 1 ? L.hi[1] : L.hi[2]);
        if (L.mode == 7u) { m = 2; sub = B.pat3[L.pat_idx - UASTC_PAT7_BASE]; }
        else {
            // BC7 subset of a texel: the BC7 subset that takes its endpoints from the texel's UASTC subset
#pragma unroll
            for (int i = 0; i < 16; i++) { const uint32_t a = (L.pattern >> (2 * i)) & 3u; sub |= (a == src0 ? 0u : (a == src1 ? 1u : 2u)) << (2 * i); }
            m = L.subsets == 3u ? 2u : (opaque ? (L.wbits == 3u ? 1u : 3u) : 7u);
        }
    }
    const uint32_t ib = m == 1u ? 3u : (m == 6u ? 4u : 2u);
    unsigned long long idx = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) idx |= (unsigned long long)B.wmap[ib - 2u][L.wbits][L.w0[i]] << (4 * i);
    bc7_pack_general(B, m, part, anc1, anc2, e0, e1, sub, idx, opaque, out);
    return true;
}

// One ETC1S block (+ optional alpha-slice block) -> BC7 mode 5.  The block's colours are base + {-b, -a, +a, +b} (clamped): the BC7
// endpoints are the darkest and brightest colours actually selected, every selector takes the nearest of the four interpolants.
UASTC_HD void etc1s_to_bc7(const Bc7Shared &B, uint32_t ep, uint32_t sel, bool has_alpha, uint32_t aep, uint32_t asel, uint32_t out[4]) {
    uint32_t col[4];
#pragma unroll
    for (int k = 0; k < 4; k++) col[k] = etc1s_color(ep, k);
    uint32_t used = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) used |= 1u << ((sel >> (2 * i)) & 3u);          // selector of texel (x, y) sits at bits 8y + 2x = 2i
    const uint32_t smin = used & 1u ? 0u : (used & 2u ? 1u : (used & 4u ? 2u : 3u)), smax = used & 8u ? 3u : (used & 4u ? 2u : (used & 2u ? 1u : 0u));
    uint32_t c0 = 0, c1 = 0, map = 0;
    if (smin == smax) {
        const uint32_t c = col[smin];
#pragma unroll
        for (uint32_t k = 0; k < 3; k++) { const uint32_t v = (c >> (8u * k)) & 255u; c0 |= (uint32_t)B.solid5[v][0] << (8u * k); c1 |= (uint32_t)B.solid5[v][1] << (8u * k); }
        map = 0x55u;          // every selector -> index 1
    } else {
        const uint32_t lc = col[smin], hc = col[smax]; uint32_t L8 = 0, H8 = 0;
#pragma unroll
        for (uint32_t k = 0; k < 3; k++) {
            const uint32_t ql = B.q7[(lc >> (8u * k)) & 255u], qh = B.q7[(hc >> (8u * k)) & 255u];
            c0 |= ql << (8u * k); c1 |= qh << (8u * k); L8 |= b7_expand(ql, 7) << (8u * k); H8 |= b7_expand(qh, 7) << (8u * k);
        }
#pragma unroll
        for (uint32_t s = 0; s < 4; s++) {
            uint32_t bestj = 0, beste = 0xffffffffu;
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                const uint32_t w = j == 0 ? 0u : (j == 1 ? 21u : (j == 2 ? 43u : 64u)); uint32_t e = 0;
                for (uint32_t k = 0; k < 3; k++) e += b7_absdiff(b7_interp((L8 >> (8u * k)) & 255u, (H8 >> (8u * k)) & 255u, w), (col[s] >> (8u * k)) & 255u);
                if (e < beste) { beste = e; bestj = j; }
            }
            map |= bestj << (2u * s);
        }
    }
    uint32_t ci = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) ci |= ((map >> (2u * ((sel >> (2 * i)) & 3u))) & 3u) << (2 * i);
    uint32_t a0 = 255, a1 = 255, ai = 0;
    if (has_alpha) {
        uint32_t av[4], aused = 0, amap = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) av[k] = (etc1s_color(aep, k) >> 8) & 255u;          // alpha = G of the alpha slice
#pragma unroll
        for (int i = 0; i < 16; i++) aused |= 1u << ((asel >> (2 * i)) & 3u);
        const uint32_t amin = aused & 1u ? 0u : (aused & 2u ? 1u : (aused & 4u ? 2u : 3u)), amax = aused & 8u ? 3u : (aused & 4u ? 2u : (aused & 2u ? 1u : 0u));
        a0 = av[amin]; a1 = av[amax];
#pragma unroll
        for (uint32_t s = 0; s < 4; s++) {
            uint32_t bestj = 0, beste = 0xffffffffu;
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) { const uint32_t w = j == 0 ? 0u : (j == 1 ? 21u : (j == 2 ? 43u : 64u)), e = b7_absdiff(b7_interp(a0, a1, w), av[s]); if (e < beste) { beste = e; bestj = j; } }
            amap |= bestj << (2u * s);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) ai |= ((amap >> (2u * ((asel >> (2 * i)) & 3u))) & 3u) << (2 * i);
    }
    bc7_pack_mode5(0, c0, c1, a0, a1, ci, ai, out);
}
