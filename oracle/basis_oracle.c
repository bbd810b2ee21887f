/*
 * basis_oracle.c -- CPU restatement of the KTX2 / Basis Universal (ETC1S "BasisLZ" and UASTC)
 * transcode-to-RGBA32 that the reference's V2 texture path runs inside its WASM worker.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * What it restates
 *   reference call site : src/lib/KTX2Loader.js:469-580 (BasisWorker.transcode: KTX2File,
 *                         startTranscoding :506, transcodeImage :551-552 per face/mip/layer,
 *                         concat(layerMips) :565), container fields per
 *                         src/lib/ktx-parse.module.js (function Pi).
 *   arithmetic          : third-party Basis Universal transcoder shipped with three 0.153.0
 *                         (src/V2/player.ts:97; yarn.lock:977-980); fixtures written by
 *                         "Basis Universal 1.16".  Source NOT in /root/reference; this file
 *                         restates the published .ktx2/BasisLZ format.
 *
 * PARITY UNPINNED against upstream binaries (no reference tests, no transcoder here).
 * Substitute pins: the reference's 50 .ktx2 fixtures with the B.4 oracles (every section and
 * every slice consumes exactly its byte length; indices in range; I-frames hold no CR block).
 * RGBA32 texels are pure integer functions of the decoded indices (B.5).
 *
 * Target format: RGBA32 only (the parity target named by north_star; the reference's own
 * fallback, KTX2Loader.js:682-687).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

int uvo_uastc_block_to_rgba(const uint8_t *blk, uint8_t *rgba /*[16][4]*/);   /* uastc_oracle.c */

/* ------------------------------------------------------------------ LSB-first bit reader */
typedef struct { const uint8_t *b; size_t n; uint64_t pos; } bits_t;
static inline uint32_t bget(bits_t *B, int n) {
    if (n == 0) return 0;
    uint64_t v = 0; size_t byte = (size_t)(B->pos >> 3);
    for (int i = 0; i < 6 && byte + i < B->n; i++) v |= (uint64_t)B->b[byte + i] << (8 * i);
    v >>= (B->pos & 7); B->pos += (uint64_t)n;
    return (uint32_t)(v & ((1ull << n) - 1));
}
static inline size_t bytes_used(const bits_t *B) { return (size_t)((B->pos + 7) >> 3); }

/* ------------------------------------------------------------------ canonical Huffman (B.2) */
#define HUFF_FAST_BITS 10
typedef struct {
    int total, used, maxl;
    uint8_t *size;                 /* [total] */
    uint16_t *sorted;              /* symbols sorted by (len, sym) */
    uint32_t first_code[18], first_idx[18], count[18];
    int16_t fast[1 << HUFF_FAST_BITS]; uint8_t fast_len[1 << HUFF_FAST_BITS];
} huff_t;

static void huff_free(huff_t *h) { free(h->size); free(h->sorted); memset(h, 0, sizeof *h); }

static int huff_build(huff_t *h, const uint8_t *sizes, int total) {
    memset(h, 0, sizeof *h);
    h->total = total; h->size = (uint8_t *)malloc((size_t)total + 1); memcpy(h->size, sizes, (size_t)total);
    for (int i = 0; i < total; i++) { if (sizes[i] > 16) return -1; if (sizes[i]) { h->count[sizes[i]]++; h->used++; if (sizes[i] > h->maxl) h->maxl = sizes[i]; } }
    h->sorted = (uint16_t *)malloc(((size_t)h->used + 1) * 2);
    uint32_t code = 0, idx = 0;
    for (int l = 1; l <= 16; l++) { code = (code + h->count[l - 1]) << 1; h->first_code[l] = code; h->first_idx[l] = idx; idx += h->count[l]; }
    uint32_t fill[18]; memcpy(fill, h->first_idx, sizeof fill);
    for (int s = 0; s < total; s++) if (sizes[s]) h->sorted[fill[sizes[s]]++] = (uint16_t)s;
    for (int i = 0; i < (1 << HUFF_FAST_BITS); i++) h->fast[i] = -1;
    for (int l = 1; l <= HUFF_FAST_BITS && l <= h->maxl; l++) {
        for (uint32_t k = 0; k < h->count[l]; k++) {
            uint32_t c = h->first_code[l] + k, rev = 0;
            for (int i = 0; i < l; i++) rev |= ((c >> i) & 1u) << (l - 1 - i);
            for (uint32_t hi = 0; hi < (1u << (HUFF_FAST_BITS - l)); hi++) {
                h->fast[rev | (hi << l)] = (int16_t)h->sorted[h->first_idx[l] + k]; h->fast_len[rev | (hi << l)] = (uint8_t)l;
            }
        }
    }
    return 0;
}
static inline int huff_dec(const huff_t *h, bits_t *B) {
    if (h->used == 0) return 0;
    /* peek */
    uint64_t v = 0; size_t byte = (size_t)(B->pos >> 3);
    for (int i = 0; i < 4 && byte + i < B->n; i++) v |= (uint64_t)B->b[byte + i] << (8 * i);
    v >>= (B->pos & 7);
    int16_t f = h->fast[v & ((1u << HUFF_FAST_BITS) - 1)];
    if (f >= 0) { B->pos += h->fast_len[v & ((1u << HUFF_FAST_BITS) - 1)]; return f; }
    uint32_t code = 0;
    for (int l = 1; l <= h->maxl; l++) {
        code = (code << 1) | (uint32_t)((v >> (l - 1)) & 1);
        if (h->count[l] && code >= h->first_code[l] && code - h->first_code[l] < h->count[l]) {
            B->pos += (uint64_t)l; return h->sorted[h->first_idx[l] + (code - h->first_code[l])];
        }
    }
    B->pos += 16; return -1;
}
static const uint8_t CL_ORDER[21] = {17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16};

static int huff_read(bits_t *B, huff_t *h) {
    memset(h, 0, sizeof *h);
    int total = (int)bget(B, 14);
    if (total == 0) return 0;
    int ncl = (int)bget(B, 5);
    if (ncl < 1 || ncl > 21) return -1;
    uint8_t cls[21]; memset(cls, 0, sizeof cls);
    for (int i = 0; i < ncl; i++) cls[CL_ORDER[i]] = (uint8_t)bget(B, 3);
    huff_t ch; if (huff_build(&ch, cls, 21)) { huff_free(&ch); return -1; }
    uint8_t *sizes = (uint8_t *)calloc((size_t)total + 160, 1); int n = 0, rc = 0;
    while (n < total) {
        int c = huff_dec(&ch, B);
        if (c < 0) { rc = -1; break; }
        if (c <= 16) sizes[n++] = (uint8_t)c;
        else if (c == 17) n += (int)bget(B, 3) + 3;
        else if (c == 18) n += (int)bget(B, 7) + 11;
        else {
            int rep = c == 19 ? (int)bget(B, 2) + 3 : (int)bget(B, 7) + 7;
            if (n == 0) { rc = -1; break; }
            uint8_t pv = sizes[n - 1];
            for (int k = 0; k < rep && n < total + 150; k++) sizes[n++] = pv;
        }
    }
    huff_free(&ch);
    if (!rc && n != total) rc = -1;
    if (!rc) rc = huff_build(h, sizes, total);
    free(sizes);
    return rc;
}
static uint32_t vlc(bits_t *B, int cb) {
    uint32_t v = 0; int ofs = 0;
    for (;;) {
        uint32_t ch = bget(B, cb + 1);
        v |= (ch & ((1u << cb) - 1)) << ofs; ofs += cb;
        if (!(ch & (1u << cb))) return v;
        if (ofs >= 32) return v;
    }
}

/* ------------------------------------------------------------------ ETC1S -> RGBA (B.5) */
static const int INTEN[8][4] = {{-8, -2, 2, 8}, {-17, -5, 5, 17}, {-29, -9, 9, 29}, {-42, -13, 13, 42},
                                {-60, -18, 18, 60}, {-80, -24, 24, 80}, {-106, -33, 33, 106}, {-183, -47, 47, 183}};
static inline uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

void uvo_ktx2_free(uvo_ktx2_image *m) {
    if (!m) return;
    free(m->rgba); free(m->endpoint_idx); free(m->selector_idx);
    memset(m, 0, sizeof *m);
}

static const uint8_t KTX2_ID[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x32, 0x30, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};

int uvo_ktx2_decode(const uint8_t *b, size_t len, uvo_ktx2_image *out) {
    memset(out, 0, sizeof *out);
    int rc = 0;
    uint8_t (*ep)[4] = NULL; uint8_t (*sel)[4] = NULL;
    huff_t epm, dem, sm, rle; memset(&epm, 0, sizeof epm); memset(&dem, 0, sizeof dem); memset(&sm, 0, sizeof sm); memset(&rle, 0, sizeof rle);
    uint16_t *rowe[2] = {NULL, NULL}; uint8_t *rowp = NULL; uint32_t *hist = NULL;
    uint16_t *alpha_e = NULL, *alpha_s = NULL;
#define FAIL(code) do { rc = (code); goto done; } while (0)
    if (len < 80 + 24 || memcmp(b, KTX2_ID, 12)) FAIL(UVO_ERR_CORRUPT);
    uint32_t vk = rd32(b + 12), w = rd32(b + 20), h = rd32(b + 24), depth = rd32(b + 28), layers = rd32(b + 32), faces = rd32(b + 36), levels = rd32(b + 40), sc = rd32(b + 44);
    uint32_t dfdOff = rd32(b + 48), dfdLen = rd32(b + 52), kvdOff = rd32(b + 56), kvdLen = rd32(b + 60);
    uint64_t sgdOff = rd64(b + 64), sgdLen = rd64(b + 72);
    if (vk != 0 || depth != 0 || faces != 1 || w == 0 || h == 0) FAIL(UVO_ERR_UNSUPPORTED);
    if (levels == 0) levels = 1;
    uint32_t nl = layers ? layers : 1;
    if (levels != 1) FAIL(UVO_ERR_UNSUPPORTED);   /* UVOL textures carry no mips (scripts/Encoder.py:290) */
    if ((uint64_t)dfdOff + dfdLen > len || (uint64_t)kvdOff + kvdLen > len || sgdOff + sgdLen > len || dfdLen < 44) FAIL(UVO_ERR_CORRUPT);
    uint64_t lvOff = rd64(b + 80), lvLen = rd64(b + 88);
    if (lvOff + lvLen > len) FAIL(UVO_ERR_TRUNCATED);
    int color_model = b[dfdOff + 12], transfer = b[dfdOff + 14], dflags = b[dfdOff + 15];
    int nsamples = (int)((rd16(b + dfdOff + 10) - 24) / 16);
    int chan0 = b[dfdOff + 28 + 3] & 0xF;
    out->width = w; out->height = h; out->layers = nl; out->levels = levels; out->faces = faces;
    out->dfd_transfer = transfer; out->dfd_flags = dflags;
    out->is_uastc = (color_model == 166);
    /* KV: any key "KTXanimData" => video */
    for (uint32_t p = kvdOff; p + 4 <= kvdOff + kvdLen;) {
        uint32_t kl = rd32(b + p); p += 4;
        if (p + kl > kvdOff + kvdLen) break;
        if (kl >= 11 && !memcmp(b + p, "KTXanimData", 11)) out->is_video = 1;
        p += (kl + 3) & ~3u;
    }
    const uint32_t bx = (w + 3) / 4, by = (h + 3) / 4, nblk = bx * by;
    out->rgba_bytes = (size_t)nl * w * h * 4;
    out->rgba = (uint8_t *)malloc(out->rgba_bytes + 64);

    if (out->is_uastc) {
        if (sc != 0) FAIL(UVO_ERR_UNSUPPORTED);    /* Zstd level supercompression not restated */
        out->has_alpha = (chan0 == 3);
        if (lvLen < (uint64_t)nl * nblk * 16) FAIL(UVO_ERR_TRUNCATED);
        for (uint32_t L = 0; L < nl; L++) {
            const uint8_t *src = b + lvOff + (size_t)L * nblk * 16; uint8_t *dst = out->rgba + (size_t)L * w * h * 4;
            for (uint32_t yb = 0; yb < by; yb++) for (uint32_t xb = 0; xb < bx; xb++) {
                uint8_t px[16][4];
                if (uvo_uastc_block_to_rgba(src + ((size_t)yb * bx + xb) * 16, &px[0][0])) FAIL(UVO_ERR_CORRUPT);
                for (uint32_t y = 0; y < 4 && yb * 4 + y < h; y++) for (uint32_t x = 0; x < 4 && xb * 4 + x < w; x++)
                    memcpy(dst + (((size_t)yb * 4 + y) * w + xb * 4 + x) * 4, px[y * 4 + x], 4);
            }
        }
        goto done;
    }
    if (color_model != 163 || sc != 1) FAIL(UVO_ERR_UNSUPPORTED);
    out->has_alpha = (nsamples == 2);
    /* ---- BasisLZ global data (B.2) */
    if (sgdLen < 20 + 20ull * nl) FAIL(UVO_ERR_CORRUPT);
    const uint8_t *g = b + sgdOff;
    uint32_t ec = rd16(g), scnt = rd16(g + 2), eb = rd32(g + 4), sb = rd32(g + 8), tb = rd32(g + 12), xb_ = rd32(g + 16);
    (void)xb_;
    if (20 + 20ull * nl + eb + sb + tb > sgdLen || ec == 0 || scnt == 0) FAIL(UVO_ERR_CORRUPT);
    const uint8_t *descs = g + 20, *epb = g + 20 + 20 * nl, *selb = epb + eb, *tabb = selb + sb;
    out->endpoint_count = ec; out->selector_count = scnt; out->endpoints_bytes = eb; out->selectors_bytes = sb; out->tables_bytes = tb;
    ep = (uint8_t (*)[4])malloc((size_t)ec * 4); sel = (uint8_t (*)[4])malloc((size_t)scnt * 4);
    { /* endpoints */
        bits_t B = {epb, eb, 0}; huff_t m0, m1, m2, mi;
        int bad = huff_read(&B, &m0) | huff_read(&B, &m1) | huff_read(&B, &m2) | huff_read(&B, &mi);
        int gray = (int)bget(&B, 1);
        int prev[3] = {16, 16, 16}, pint = 0;
        for (uint32_t i = 0; i < ec && !bad; i++) {
            int d = huff_dec(&mi, &B); if (d < 0) { bad = 1; break; }
            pint = (pint + d) & 7;
            for (int c = 0; c < (gray ? 1 : 3); c++) {
                const huff_t *m = prev[c] <= 9 ? &m0 : (prev[c] <= 21 ? &m1 : &m2);
                d = huff_dec(m, &B); if (d < 0) { bad = 1; break; }
                prev[c] = (prev[c] + d) & 31;
            }
            if (gray) prev[1] = prev[2] = prev[0];
            ep[i][0] = (uint8_t)prev[0]; ep[i][1] = (uint8_t)prev[1]; ep[i][2] = (uint8_t)prev[2]; ep[i][3] = (uint8_t)pint;
        }
        out->endpoints_used = (uint32_t)bytes_used(&B);
        huff_free(&m0); huff_free(&m1); huff_free(&m2); huff_free(&mi);
        if (bad) FAIL(UVO_ERR_CORRUPT);
    }
    { /* selectors */
        bits_t B = {selb, sb, 0};
        int glob = (int)bget(&B, 1), hyb = (int)bget(&B, 1), raw = (int)bget(&B, 1);
        if (glob || hyb) FAIL(UVO_ERR_UNSUPPORTED);
        if (raw) { for (uint32_t i = 0; i < scnt; i++) for (int j = 0; j < 4; j++) sel[i][j] = (uint8_t)bget(&B, 8); }
        else {
            huff_t dm; if (huff_read(&B, &dm)) { huff_free(&dm); FAIL(UVO_ERR_CORRUPT); }
            uint8_t pb[4] = {0, 0, 0, 0}; int bad = 0;
            for (uint32_t i = 0; i < scnt && !bad; i++) {
                for (int j = 0; j < 4; j++) {
                    if (i == 0) pb[j] = (uint8_t)bget(&B, 8);
                    else { int d = huff_dec(&dm, &B); if (d < 0) { bad = 1; break; } pb[j] ^= (uint8_t)d; }
                    sel[i][j] = pb[j];
                }
            }
            huff_free(&dm);
            if (bad) FAIL(UVO_ERR_CORRUPT);
        }
        out->selectors_used = (uint32_t)bytes_used(&B);
    }
    uint32_t hsize;
    { /* tables */
        bits_t B = {tabb, tb, 0};
        if (huff_read(&B, &epm) | huff_read(&B, &dem) | huff_read(&B, &sm) | huff_read(&B, &rle)) FAIL(UVO_ERR_CORRUPT);
        hsize = bget(&B, 13);
        out->tables_used = (uint32_t)bytes_used(&B);
        if (hsize == 0) FAIL(UVO_ERR_CORRUPT);
    }
    /* ---- slices (B.3) */
    out->endpoint_idx = (uint16_t *)calloc((size_t)nl * nblk, 2); out->selector_idx = (uint16_t *)calloc((size_t)nl * nblk, 2);
    if (out->has_alpha) { alpha_e = (uint16_t *)calloc((size_t)nl * nblk, 2); alpha_s = (uint16_t *)calloc((size_t)nl * nblk, 2); }
    rowe[0] = (uint16_t *)calloc(bx + 1, 2); rowe[1] = (uint16_t *)calloc(bx + 1, 2); rowp = (uint8_t *)calloc(bx + 1, 1);
    hist = (uint32_t *)calloc(hsize, 4);
    for (int pass = 0; pass < (out->has_alpha ? 2 : 1); pass++) {
        for (uint32_t L = 0; L < nl; L++) {
            const uint8_t *d = descs + 20 * L;
            uint32_t off = rd32(d + 4 + 8 * pass), ln = rd32(d + 8 + 8 * pass);
            if ((uint64_t)off + ln > lvLen) FAIL(UVO_ERR_TRUNCATED);
            bits_t B = {b + lvOff + off, ln, 0};
            uint16_t *E = (pass ? alpha_e : out->endpoint_idx) + (size_t)L * nblk, *S = (pass ? alpha_s : out->selector_idx) + (size_t)L * nblk;
            const uint16_t *PE = L ? E - nblk : NULL, *PS = L ? S - nblk : NULL;
            memset(hist, 0, (size_t)hsize * 4); memset(rowe[0], 0, (bx + 1) * 2); memset(rowe[1], 0, (bx + 1) * 2); memset(rowp, 0, bx + 1);
            uint32_t rover = hsize / 2, rle_cnt = 0, prev_sym = 0, rep = 0, prev_ep = 0, bits = 0;
            for (uint32_t y = 0; y < by; y++) {
                int cur = y & 1;
                for (uint32_t x = 0; x < bx; x++) {
                    if ((x & 1) == 0) {
                        if ((y & 1) == 0) {
                            if (rep) { rep--; bits = prev_sym; }
                            else {
                                int sy = huff_dec(&epm, &B); if (sy < 0) FAIL(UVO_ERR_CORRUPT);
                                bits = (uint32_t)sy;
                                if (bits == 256) { rep = vlc(&B, 4) + 3 - 1; bits = prev_sym; } else prev_sym = bits;
                            }
                            rowp[x] = (uint8_t)(bits >> 4);
                        } else bits = rowp[x];
                    }
                    uint32_t pred = bits & 3; bits >>= 2;
                    if (pass == 0) out->pred_hist[pred]++;
                    uint32_t e, s = 0; int cr = 0;
                    if (pred == 0) { if (x == 0) FAIL(UVO_ERR_CORRUPT); e = prev_ep; }
                    else if (pred == 1) { if (y == 0) FAIL(UVO_ERR_CORRUPT); e = rowe[cur ^ 1][x]; }
                    else if (pred == 2) {
                        if (out->is_video) { if (!PE) FAIL(UVO_ERR_CORRUPT); e = PE[y * bx + x]; s = PS[y * bx + x]; cr = 1; }
                        else { if (x == 0 || y == 0) FAIL(UVO_ERR_CORRUPT); e = rowe[cur ^ 1][x - 1]; }
                    } else {
                        int dsy = huff_dec(&dem, &B); if (dsy < 0) FAIL(UVO_ERR_CORRUPT);
                        e = (uint32_t)dsy + prev_ep; if (e >= ec) e -= ec;
                    }
                    rowe[cur][x] = (uint16_t)e; prev_ep = e;
                    if (!cr) {
                        if (rle_cnt > 0) { rle_cnt--; s = hist[0]; }
                        else {
                            int sy = huff_dec(&sm, &B); if (sy < 0) FAIL(UVO_ERR_CORRUPT);
                            s = (uint32_t)sy;
                            if (s == scnt + hsize) {
                                int rr = huff_dec(&rle, &B); if (rr < 0) FAIL(UVO_ERR_CORRUPT);
                                rle_cnt = (rr == 63) ? vlc(&B, 7) + 3 : (uint32_t)rr + 3;
                                s = hist[0]; rle_cnt--;
                            } else if (s >= scnt) {
                                uint32_t i = s - scnt; if (i >= hsize) FAIL(UVO_ERR_CORRUPT);
                                s = hist[i];
                                if (i) { uint32_t tmp = hist[i / 2]; hist[i / 2] = hist[i]; hist[i] = tmp; }
                            } else { hist[rover++] = s; if (rover == hsize) rover = hsize / 2; }
                        }
                    }
                    if (e >= ec || s >= scnt) FAIL(UVO_ERR_CORRUPT);
                    E[y * bx + x] = (uint16_t)e; S[y * bx + x] = (uint16_t)s;
                }
            }
            out->slices++;
            if (bytes_used(&B) == ln) out->slices_exact++;
        }
    }
    /* ---- blocks -> RGBA32 (B.5); layers concatenated layer-major (KTX2Loader.js:565) */
    for (uint32_t L = 0; L < nl; L++) {
        uint8_t *dst = out->rgba + (size_t)L * w * h * 4;
        for (uint32_t yb = 0; yb < by; yb++) for (uint32_t xb = 0; xb < bx; xb++) {
            size_t bi = (size_t)L * nblk + (size_t)yb * bx + xb;
            const uint8_t *E = ep[out->endpoint_idx[bi]], *S = sel[out->selector_idx[bi]];
            uint8_t col[4][3];
            for (int k = 0; k < 4; k++) for (int c = 0; c < 3; c++) col[k][c] = clamp255((((int)E[c] << 3) | (E[c] >> 2)) + INTEN[E[3]][k]);
            uint8_t acol[4] = {255, 255, 255, 255}; const uint8_t *AS = NULL;
            if (alpha_e) { const uint8_t *AE = ep[alpha_e[bi]]; AS = sel[alpha_s[bi]]; for (int k = 0; k < 4; k++) acol[k] = clamp255((((int)AE[1] << 3) | (AE[1] >> 2)) + INTEN[AE[3]][k]); }
            for (uint32_t y = 0; y < 4 && yb * 4 + y < h; y++) for (uint32_t x = 0; x < 4 && xb * 4 + x < w; x++) {
                uint8_t *o = dst + (((size_t)yb * 4 + y) * w + xb * 4 + x) * 4;
                int k = (S[y] >> (2 * x)) & 3;
                o[0] = col[k][0]; o[1] = col[k][1]; o[2] = col[k][2];
                o[3] = AS ? acol[(AS[y] >> (2 * x)) & 3] : 255;
            }
        }
    }
done:
    free(alpha_e); free(alpha_s);
    free(ep); free(sel); huff_free(&epm); huff_free(&dem); huff_free(&sm); huff_free(&rle);
    free(rowe[0]); free(rowe[1]); free(rowp); free(hist);
    out->status = rc;
    if (rc) { int s = rc; uvo_ktx2_free(out); out->status = s; }
    return rc;
#undef FAIL
}
