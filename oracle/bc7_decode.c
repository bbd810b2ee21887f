/*
 * bc7_decode.c -- BC7 (BPTC) block decoder, all eight modes.  TEST INFRASTRUCTURE (see oracle.h): the independent check of the
 * product's BC7 target format (csrc/bc7_core.h).  Written from the published BC7 / BPTC block layout (Khronos Data Format
 * Specification, "BPTC compressed texture image formats"): mode = number of low zero bits; per mode the subset count, partition
 * bits, rotation / index-selection bits, colour / alpha endpoint precision, p-bits and index widths below; endpoints are stored
 * channel by channel; the anchor index of every subset omits its top bit; interpolation = ((64 - w) * e0 + w * e1 + 32) >> 6 with
 * the 2- / 3- / 4-bit weight tables.
 */
#include <string.h>
#include "oracle.h"
#include "bc7_tables.h"

typedef struct { int ns, pb, rb, isb, cb, ab, epb, spb, ib, ib2; } Bc7Mode;
static const Bc7Mode MODES[8] = {
    {3, 4, 0, 0, 4, 0, 1, 0, 3, 0}, {2, 6, 0, 0, 6, 0, 0, 1, 3, 0}, {3, 6, 0, 0, 5, 0, 0, 0, 2, 0}, {2, 6, 0, 0, 7, 0, 1, 0, 2, 0},
    {1, 0, 2, 1, 5, 6, 0, 0, 2, 3}, {1, 0, 2, 0, 7, 8, 0, 0, 2, 2}, {1, 0, 0, 0, 7, 7, 1, 0, 4, 0}, {2, 6, 0, 0, 5, 5, 1, 0, 2, 0}};
static const int W2[4] = {0, 21, 43, 64}, W3[8] = {0, 9, 18, 27, 37, 46, 55, 64}, W4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};

typedef struct { const uint8_t *p; int pos; } Rd;
static uint32_t rd(Rd *r, int n) { uint32_t v = 0; for (int i = 0; i < n; i++, r->pos++) v |= (uint32_t)((r->p[r->pos >> 3] >> (r->pos & 7)) & 1) << i; return v; }
static int weight(int bits, int i) { return bits == 2 ? W2[i] : (bits == 3 ? W3[i] : W4[i]); }

/* block: 16 bytes -> out: 16 texels x RGBA8 (raster order).  Returns 0, or -1 for the reserved mode (all zero output). */
int uvo_bc7_decode_block(const uint8_t *block, uint8_t *out) {
    int mode = 0;
    while (mode < 8 && !((block[0] >> mode) & 1)) mode++;
    if (mode == 8) { memset(out, 0, 64); return -1; }
    const Bc7Mode M = MODES[mode];
    Rd r = {block, mode + 1};
    const int part = (int)rd(&r, M.pb), rot = (int)rd(&r, M.rb), isel = (int)rd(&r, M.isb);
    int ep[6][4];
    for (int c = 0; c < 3; c++) for (int e = 0; e < 2 * M.ns; e++) ep[e][c] = (int)rd(&r, M.cb);
    for (int e = 0; e < 2 * M.ns; e++) ep[e][3] = M.ab ? (int)rd(&r, M.ab) : 255;
    int cbits = M.cb, abits = M.ab;
    if (M.epb) { for (int e = 0; e < 2 * M.ns; e++) { const int p = (int)rd(&r, 1); for (int c = 0; c < 3; c++) ep[e][c] = (ep[e][c] << 1) | p; if (M.ab) ep[e][3] = (ep[e][3] << 1) | p; } cbits++; if (abits) abits++; }
    if (M.spb) { for (int s = 0; s < M.ns; s++) { const int p = (int)rd(&r, 1); for (int e = 2 * s; e < 2 * s + 2; e++) for (int c = 0; c < 3; c++) ep[e][c] = (ep[e][c] << 1) | p; } cbits++; }
    for (int e = 0; e < 2 * M.ns; e++) {
        for (int c = 0; c < 3; c++) { ep[e][c] <<= (8 - cbits); ep[e][c] |= ep[e][c] >> cbits; }
        if (M.ab) { ep[e][3] <<= (8 - abits); ep[e][3] |= ep[e][3] >> abits; }
    }
    int subset[16], anchor[3] = {0, 0, 0};
    for (int i = 0; i < 16; i++) subset[i] = M.ns == 1 ? 0 : (M.ns == 2 ? BC7_PART2[part][i] : BC7_PART3[part][i]);
    if (M.ns == 2) anchor[1] = BC7_ANCHOR2[part];
    if (M.ns == 3) { anchor[1] = BC7_ANCHOR3A[part]; anchor[2] = BC7_ANCHOR3B[part]; }
    int idx[16], idx2[16];
    for (int i = 0; i < 16; i++) { const int a = i == anchor[subset[i]]; idx[i] = (int)rd(&r, M.ib - a); }
    for (int i = 0; i < 16; i++) idx2[i] = M.ib2 ? (int)rd(&r, M.ib2 - (i == 0)) : 0;
    if (r.pos != 128) { memset(out, 0, 64); return -1; }
    for (int i = 0; i < 16; i++) {
        const int *e0 = ep[2 * subset[i]], *e1 = ep[2 * subset[i] + 1];
        int cw, aw;
        if (!M.ib2) cw = aw = weight(M.ib, idx[i]);
        else if (isel) { cw = weight(M.ib2, idx2[i]); aw = weight(M.ib, idx[i]); }
        else { cw = weight(M.ib, idx[i]); aw = weight(M.ib2, idx2[i]); }
        int px[4];
        for (int c = 0; c < 3; c++) px[c] = ((64 - cw) * e0[c] + cw * e1[c] + 32) >> 6;
        px[3] = ((64 - aw) * e0[3] + aw * e1[3] + 32) >> 6;
        if (rot) { const int t = px[3]; px[3] = px[rot - 1]; px[rot - 1] = t; }
        for (int c = 0; c < 4; c++) out[4 * i + c] = (uint8_t)px[c];
    }
    return 0;
}

/* blocks in block-raster order (bx x by) -> RGBA8 image w x h (ragged edges cropped) */
int uvo_bc7_decode_image(const uint8_t *blocks, uint32_t w, uint32_t h, uint8_t *rgba) {
    const uint32_t bx = (w + 3) / 4, by = (h + 3) / 4; int bad = 0;
    for (uint32_t y = 0; y < by; y++) for (uint32_t x = 0; x < bx; x++) {
        uint8_t t[64];
        if (uvo_bc7_decode_block(blocks + 16 * ((size_t)y * bx + x), t)) bad++;
        for (uint32_t py = 0; py < 4 && 4 * y + py < h; py++) for (uint32_t pxx = 0; pxx < 4 && 4 * x + pxx < w; pxx++)
            memcpy(rgba + 4 * ((size_t)(4 * y + py) * w + 4 * x + pxx), t + 4 * (4 * py + pxx), 4);
    }
    return bad;
}
