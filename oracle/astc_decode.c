/* astc_decode.c -- ASTC LDR decoder for 4x4 blocks: the INDEPENDENT check of the product's UVOL_TEX_ASTC_4x4 target.
 *
 * TEST INFRASTRUCTURE (see oracle.h): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Written from the ASTC specification (Khronos data format specification, "ASTC compressed texture image formats"), not from
 * the product's packer: the block is taken apart the way hardware does it --
 *   block mode -> weight grid, weight range, plane count;  partition count / seed / colour endpoint mode;
 *   the endpoint range is NOT stored: it is the largest range whose integer-sequence encoding fits the bits that remain
 *   between the header and the (reversed) weight field;  trits / quints are unbundled from their 8- / 7-bit blocks by the
 *   spec's bit equations;  RGB / RGBA direct endpoints apply blue contraction when the second sum is smaller;
 *   the partition of a texel is the spec's hash (small-block variant);  weights interpolate 16-bit expanded endpoints and the
 *   top byte is the UNORM8 result.
 * Scope: what a UASTC-derived block can be -- 4x4 weight grid, 1..3 partitions with one shared endpoint mode, modes 4 (LA direct),
 * 8 (RGB direct), 12 (RGBA direct), bit-only weight ranges, void-extent blocks.  Anything else is reported as a bad block.
 * The reference selects this target for UASTC sources on GPUs with ASTC support (src/lib/KTX2Loader.js:592-600).
 */
#include <string.h>
#include "oracle.h"

static uint32_t get_bits(const uint8_t *b, uint32_t pos, uint32_t n) {      /* n <= 16, LSB first */
    uint32_t v = 0;
    for (uint32_t i = 0; i < n; i++) { const uint32_t p = pos + i; if (p < 128) v |= (uint32_t)((b[p >> 3] >> (p & 7)) & 1) << i; }
    return v;
}

static uint32_t hash52(uint32_t p) {
    p ^= p >> 15; p -= p << 17; p += p << 7; p += p << 4; p ^= p >> 5; p += p << 16; p ^= p >> 7; p ^= p >> 3; p ^= p << 6; p ^= p >> 17;
    return p;
}
static int select_partition(int seed, int x, int y, int count) {              /* 4x4 footprint: "small block", z = 0 */
    x <<= 1; y <<= 1;
    seed += (count - 1) * 1024;
    const uint32_t r = hash52((uint32_t)seed);
    uint32_t s[8];
    for (int i = 0; i < 8; i++) { s[i] = (r >> (4 * i)) & 15; s[i] *= s[i]; }
    /* seeds 9..12 belong to z and do not matter at z = 0 */
    int sh1, sh2;
    if (seed & 1) { sh1 = seed & 2 ? 4 : 5; sh2 = count == 3 ? 6 : 5; }
    else { sh1 = count == 3 ? 6 : 5; sh2 = seed & 2 ? 4 : 5; }
    for (int i = 0; i < 8; i++) s[i] >>= (i & 1) ? sh2 : sh1;
    int a = (int)((s[0] * x + s[1] * y + (r >> 14)) & 63);
    int b = (int)((s[2] * x + s[3] * y + (r >> 10)) & 63);
    int c = (int)((s[4] * x + s[5] * y + (r >> 6)) & 63);
    int d = (int)((s[6] * x + s[7] * y + (r >> 2)) & 63);
    if (count < 4) d = 0;
    if (count < 3) c = 0;
    if (a >= b && a >= c && a >= d) return 0;
    if (b >= c && b >= d) return 1;
    if (c >= d) return 2;
    return 3;
}

/* the 21 endpoint ranges: {bits, trits, quints} */
static const uint8_t RANGE[21][3] = {{1, 0, 0}, {0, 1, 0}, {2, 0, 0}, {0, 0, 1}, {1, 1, 0}, {3, 0, 0}, {1, 0, 1}, {2, 1, 0}, {4, 0, 0}, {2, 0, 1}, {3, 1, 0},
                                     {5, 0, 0}, {3, 0, 1}, {4, 1, 0}, {6, 0, 0}, {4, 0, 1}, {5, 1, 0}, {7, 0, 0}, {5, 0, 1}, {6, 1, 0}, {8, 0, 0}};
static uint32_t ise_bits(uint32_t n, const uint8_t r[3]) { return n * r[0] + (r[1] ? (8 * n + 4) / 5 : 0) + (r[2] ? (7 * n + 2) / 3 : 0); }

static void trits_of(uint32_t T, uint32_t t[5]) {
#define TB(i) ((T >> (i)) & 1u)
    uint32_t C;
    if (((T >> 2) & 7u) == 7u) { C = ((T >> 5) << 2) | (T & 3u); t[4] = t[3] = 2; }
    else { C = T & 31u; if (((T >> 5) & 3u) == 3u) { t[4] = 2; t[3] = TB(7); } else { t[4] = TB(7); t[3] = (T >> 5) & 3u; } }
#define CB(i) ((C >> (i)) & 1u)
    if ((C & 3u) == 3u) { t[2] = 2; t[1] = CB(4); t[0] = (CB(3) << 1) | (CB(2) & ~CB(3) & 1u); }
    else if (((C >> 2) & 3u) == 3u) { t[2] = 2; t[1] = 2; t[0] = C & 3u; }
    else { t[2] = CB(4); t[1] = (C >> 2) & 3u; t[0] = (CB(1) << 1) | (CB(0) & ~CB(1) & 1u); }
#undef TB
}
static void quints_of(uint32_t Q, uint32_t q[3]) {
#define QB(i) ((Q >> (i)) & 1u)
    uint32_t C;
    if (((Q >> 1) & 3u) == 3u && ((Q >> 5) & 3u) == 0u) { q[2] = (QB(0) << 2) | ((QB(4) & ~QB(0) & 1u) << 1) | (QB(3) & ~QB(0) & 1u); q[1] = q[0] = 4; return; }
    if (((Q >> 1) & 3u) == 3u) { q[2] = 4; C = (((Q >> 3) & 3u) << 3) | ((~(Q >> 5) & 3u) << 1) | QB(0); }
    else { q[2] = (Q >> 5) & 3u; C = Q & 31u; }
    if ((C & 7u) == 5u) { q[1] = 4; q[0] = (C >> 3) & 3u; } else { q[1] = (C >> 3) & 3u; q[0] = C & 7u; }
#undef QB
#undef CB
}

/* integer sequence decoding of n values starting at bit `pos`: out[i] = bits | digit << nbits */
static void ise_decode(const uint8_t *blk, uint32_t pos, uint32_t n, const uint8_t r[3], uint32_t *out) {
    const uint32_t nb = r[0];
    if (r[1]) {
        for (uint32_t i = 0; i < n; i += 5) {
            uint32_t m[5], T = 0, t[5];
            static const uint8_t tw[5] = {2, 2, 1, 2, 1}, ts[5] = {0, 2, 4, 5, 7};
            for (uint32_t k = 0; k < 5; k++) {          /* absent values of a last partial bundle read as zero bits */
                if (i + k < n) { m[k] = get_bits(blk, pos, nb); pos += nb; T |= get_bits(blk, pos, tw[k]) << ts[k]; pos += tw[k]; } else m[k] = 0;
            }
            trits_of(T, t);
            for (uint32_t k = 0; k < 5 && i + k < n; k++) out[i + k] = m[k] | (t[k] << nb);
        }
    } else if (r[2]) {
        for (uint32_t i = 0; i < n; i += 3) {
            uint32_t m[3], Q = 0, q[3];
            static const uint8_t qw[3] = {3, 2, 2}, qs[3] = {0, 3, 5};
            for (uint32_t k = 0; k < 3; k++) {
                if (i + k < n) { m[k] = get_bits(blk, pos, nb); pos += nb; Q |= get_bits(blk, pos, qw[k]) << qs[k]; pos += qw[k]; } else m[k] = 0;
            }
            quints_of(Q, q);
            for (uint32_t k = 0; k < 3 && i + k < n; k++) out[i + k] = m[k] | (q[k] << nb);
        }
    } else {
        for (uint32_t i = 0; i < n; i++) { out[i] = get_bits(blk, pos, nb); pos += nb; }
    }
}

/* endpoint unquantisation to 0..255 (spec: "endpoint unquantization") */
static uint32_t unquant_endpoint(uint32_t val, const uint8_t r[3]) {
    const uint32_t bits = r[0], lo = val & ((1u << bits) - 1u), D = val >> bits;
    if (!r[1] && !r[2]) {
        uint32_t v = lo << (8 - bits), out = v, sh = bits;
        while (sh < 8) { out |= v >> sh; sh += bits; }
        return out & 255u;
    }
    const uint32_t A = (lo & 1u) ? 511u : 0u, x = lo >> 1;      /* x = the bits above bit 0: b, cb, dcb ... */
    uint32_t B = 0, C = 0;
    if (r[1]) {
        switch (bits) {
        case 1: C = 204; B = 0; break;
        case 2: C = 93; B = (x << 8) | (x << 4) | (x << 2) | (x << 1); break;              /* b000b0bb0 */
        case 3: C = 44; B = (x << 7) | (x << 2) | x; break;                                  /* cb000cbcb */
        case 4: C = 22; B = (x << 6) | x; break;                                             /* dcb000dcb */
        case 5: C = 11; B = (x << 5) | (x >> 2); break;                                      /* edcb000ed */
        default: C = 5; B = (x << 4) | (x >> 4); break;                                      /* fedcb000f */
        }
    } else {
        switch (bits) {
        case 1: C = 113; B = 0; break;
        case 2: C = 54; B = (x << 8) | (x << 3) | (x << 2); break;                           /* b0000bb00 */
        case 3: C = 26; B = (x << 7) | (x << 1) | (x >> 1); break;                           /* cb0000cbc */
        case 4: C = 13; B = (x << 6) | (x >> 1); break;                                      /* dcb0000dc */
        default: C = 6; B = (x << 5) | (x >> 3); break;                                      /* edcb0000e */
        }
    }
    uint32_t T = D * C + B;
    T ^= A;
    return (A & 0x80u) | (T >> 2);
}

static void blue_contract(int *r, int *g, int *b) { *r = (*r + *b) >> 1; *g = (*g + *b) >> 1; }

/* One block -> 16 RGBA texels (row-major).  0: decoded; 1: not a block this decoder accepts (texels are set to magenta). */
int uvo_astc_decode_block(const uint8_t *blk, uint8_t *rgba64) {
    for (int i = 0; i < 16; i++) { rgba64[4 * i] = 255; rgba64[4 * i + 1] = 0; rgba64[4 * i + 2] = 255; rgba64[4 * i + 3] = 255; }
    const uint32_t bm = get_bits(blk, 0, 11);
    if ((bm & 0x1FFu) == 0x1FCu) {                       /* void extent */
        if (bm & 0x200u) return 1;                       /* HDR */
        if (get_bits(blk, 10, 2) != 3) return 1;
        /* extent coordinates: all ones = none */
        for (uint32_t p = 12; p < 64; p += 13) if (get_bits(blk, p, 13) != 0x1FFFu) return 1;
        for (int i = 0; i < 16; i++) for (int c = 0; c < 4; c++) rgba64[4 * i + c] = (uint8_t)(get_bits(blk, 64 + 16 * c, 16) >> 8);
        return 0;
    }
    if ((bm & 3u) == 0) return 1;                        /* the 12 / 6 / 10-wide layouts and the reserved ones */
    /* first table of block modes: R1 R2 in bits 0-1, layout in bits 2-3 */
    const uint32_t R = ((bm >> 4) & 1u) | ((bm & 3u) << 1), A = (bm >> 5) & 3u, B = (bm >> 7) & 3u, H = (bm >> 9) & 1u, D = (bm >> 10) & 1u;
    uint32_t W, Hh;
    switch ((bm >> 2) & 3u) {
    case 0: W = B + 4; Hh = A + 2; break;
    case 1: W = B + 8; Hh = A + 2; break;
    case 2: W = A + 2; Hh = B + 8; break;
    default: if (B & 2u) { W = (B & 1u) + 2; Hh = A + 2; } else { W = A + 2; Hh = (B & 1u) + 6; } break;
    }
    if (W != 4 || Hh != 4) return 1;
    if (R < 2) return 1;                                  /* reserved */
    /* weight range: bits only for {2, 4, 8, 16, 32} levels; trit / quint weight ranges are outside this decoder's scope */
    static const int8_t wbits_lo[8] = {-1, -1, 1, -1, 2, -1, -1, 3}, wbits_hi[8] = {-1, -1, -1, -1, 4, -1, -1, 5};
    const int wb = H ? wbits_hi[R] : wbits_lo[R];
    if (wb < 0) return 1;
    const uint32_t planes = D + 1, parts = get_bits(blk, 11, 2) + 1;
    if (parts > 3 || (parts == 4 && D)) return 1;
    uint32_t cem, seed = 0, pos;
    if (parts == 1) { cem = get_bits(blk, 13, 4); pos = 17; }
    else {
        seed = get_bits(blk, 13, 10);
        const uint32_t c6 = get_bits(blk, 23, 6);
        if (c6 & 3u) return 1;                            /* per-partition endpoint modes */
        cem = c6 >> 2; pos = 29;
    }
    if (cem != 4 && cem != 8 && cem != 12) return 1;
    const uint32_t wtotal = 16u * planes * (uint32_t)wb;
    if (wtotal > 96 || wtotal < 24) return 1;
    const uint32_t nvals = parts * (cem == 4 ? 4u : (cem == 8 ? 6u : 8u));
    if (nvals > 18) return 1;
    const int remaining = 128 - (int)wtotal - (int)pos - (D ? 2 : 0);
    int range = -1;
    for (int r = 20; r >= 0; r--) if ((int)ise_bits(nvals, RANGE[r]) <= remaining) { range = r; break; }
    if (range < 4) return 1;                              /* fewer than 6 levels: illegal */
    uint32_t raw[18], v[18];
    ise_decode(blk, pos, nvals, RANGE[range], raw);
    for (uint32_t i = 0; i < nvals; i++) v[i] = unquant_endpoint(raw[i], RANGE[range]);
    int e0[3][4], e1[3][4];
    for (uint32_t p = 0; p < parts; p++) {
        const uint32_t *q = v + p * (nvals / parts);
        if (cem == 4) { e0[p][0] = e0[p][1] = e0[p][2] = (int)q[0]; e1[p][0] = e1[p][1] = e1[p][2] = (int)q[1]; e0[p][3] = (int)q[2]; e1[p][3] = (int)q[3]; }
        else {
            const int a0 = cem == 12 ? (int)q[6] : 255, a1 = cem == 12 ? (int)q[7] : 255;
            if (q[1] + q[3] + q[5] >= q[0] + q[2] + q[4]) {
                e0[p][0] = (int)q[0]; e0[p][1] = (int)q[2]; e0[p][2] = (int)q[4]; e0[p][3] = a0;
                e1[p][0] = (int)q[1]; e1[p][1] = (int)q[3]; e1[p][2] = (int)q[5]; e1[p][3] = a1;
            } else {
                e0[p][0] = (int)q[1]; e0[p][1] = (int)q[3]; e0[p][2] = (int)q[5]; e0[p][3] = a1; blue_contract(&e0[p][0], &e0[p][1], &e0[p][2]);
                e1[p][0] = (int)q[0]; e1[p][1] = (int)q[2]; e1[p][2] = (int)q[4]; e1[p][3] = a0; blue_contract(&e1[p][0], &e1[p][1], &e1[p][2]);
            }
        }
    }
    const uint32_t ccs = D ? get_bits(blk, 128 - wtotal - 2, 2) : 4;
    for (uint32_t i = 0; i < 16; i++) {
        const int p = parts == 1 ? 0 : select_partition((int)seed, (int)(i & 3), (int)(i >> 2), (int)parts);
        uint32_t w[2];
        for (uint32_t pl = 0; pl < planes; pl++) {
            uint32_t q = 0;
            for (int k = 0; k < wb; k++) { const uint32_t sp = (i * planes + pl) * (uint32_t)wb + (uint32_t)k; q |= get_bits(blk, 127 - sp, 1) << k; }
            /* weight unquantisation, bit-only ranges: replicate to 6 bits, then 33..63 -> +1 */
            uint32_t u;
            switch (wb) {
            case 1: u = q ? 63 : 0; break;
            case 2: u = q | (q << 2) | (q << 4); break;
            case 3: u = q | (q << 3); break;
            case 4: u = (q >> 2) | (q << 2); break;
            default: u = (q >> 4) | (q << 1); break;
            }
            if (u > 32) u++;
            w[pl] = u;
        }
        for (uint32_t c = 0; c < 4; c++) {
            const uint32_t ww = (D && c == ccs) ? w[1] : w[0];
            const uint32_t c0 = (uint32_t)e0[p][c] * 257u, c1 = (uint32_t)e1[p][c] * 257u;
            rgba64[4 * i + c] = (uint8_t)(((c0 * (64 - ww) + c1 * ww + 32) >> 6) >> 8);
        }
    }
    return 0;
}

/* blocks in block-raster order -> u8[h][w][4]; returns the number of bad blocks */
int uvo_astc_decode_image(const uint8_t *blocks, uint32_t w, uint32_t h, uint8_t *rgba) {
    const uint32_t bx = (w + 3) / 4, by = (h + 3) / 4; int bad = 0;
    for (uint32_t yb = 0; yb < by; yb++) for (uint32_t xb = 0; xb < bx; xb++) {
        uint8_t px[64];
        bad += uvo_astc_decode_block(blocks + ((size_t)yb * bx + xb) * 16, px);
        for (uint32_t y = 0; y < 4 && yb * 4 + y < h; y++) for (uint32_t x = 0; x < 4 && xb * 4 + x < w; x++)
            memcpy(rgba + (((size_t)yb * 4 + y) * w + xb * 4 + x) * 4, px + 4 * (4 * y + x), 4);
    }
    return bad;
}
