// basis_core.h -- per-unit device logic of the KTX2 / BasisLZ (ETC1S) transcode to RGBA32.
// __host__ __device__ so tests/tools/basis_emu.cpp can check the logic on the host (debug harness
// only; libuvol_b200.so has no host transcode).  Format: SURVEY.md Appendix B; reference call
// site src/lib/KTX2Loader.js:469-580 (startTranscoding :506, transcodeImage :551-552).
#pragma once
#include "uvol_internal.h"

#if defined(__CUDACC__)
#define UVOL_HD __host__ __device__ __forceinline__
#else
#define UVOL_HD static inline
#endif

// ---------------------------------------------------------------------------------------------
// LSB-first bit reader over 32-bit aligned words, bounded by the section's byte length: past the last word that holds a
// section byte it yields zeros (a corrupt stream is then caught by the callers' `consumed` check instead of walking out of the
// blob).  The batch blob is padded, so the aligned words that straddle the section's ends are always readable.  The word behind
// the one in use is requested at the previous refill, so a refill never waits for memory on the symbol chain.
struct BitRd { const uint32_t *w; uint32_t wi, lim, nxt; uint64_t buf; int nbits; uint64_t consumed; };      // nxt = word wi, requested one refill ahead of its use
UVOL_HD void br_init(BitRd &b, const uint8_t *p, uint32_t nbytes) {
    const uintptr_t a = (uintptr_t)p; const unsigned mis = (unsigned)(a & 3);
    b.w = (const uint32_t *)(a - mis); b.lim = (mis + nbytes + 3) / 4;
    b.buf = b.lim ? (uint64_t)b.w[0] >> (8 * mis) : 0; b.nbits = 32 - 8 * (int)mis; b.wi = 1; b.consumed = 0;
    b.nxt = b.lim > 1 ? b.w[1] : 0u;
}
UVOL_HD void br_refill(BitRd &b) { if (b.nbits <= 32) { b.buf |= (uint64_t)b.nxt << b.nbits; b.nbits += 32; b.wi++; b.nxt = b.wi < b.lim ? b.w[b.wi] : 0u; } }
UVOL_HD uint32_t br_peek(BitRd &b) { br_refill(b); return (uint32_t)b.buf; }
UVOL_HD void br_skip(BitRd &b, int n) { b.buf >>= n; b.nbits -= n; b.consumed += (uint64_t)n; }
UVOL_HD uint32_t br_get(BitRd &b, int n) { if (n == 0) return 0; br_refill(b); uint32_t v = (uint32_t)b.buf & ((n >= 32) ? 0xffffffffu : ((1u << n) - 1u)); br_skip(b, n); return v; }
UVOL_HD uint32_t br_vlc(BitRd &b, int cb) {
    uint32_t v = 0; int ofs = 0;
    for (;;) { uint32_t ch = br_get(b, cb + 1); v |= (ch & ((1u << cb) - 1u)) << ofs; ofs += cb; if (!(ch & (1u << cb)) || ofs >= 32) return v; }
}

// ---------------------------------------------------------------------------------------------
// Canonical Huffman (B.2).  Codes are stored MSB-first canonical, read LSB-first from the stream.
UVOL_HD int huff_decode(const HuffTable &T, const uint16_t *sorted_pool, BitRd &b) {
    if (T.used == 0) return 0;
    const uint32_t v = br_peek(b);
    const uint32_t e = T.fast[v & ((1u << UVOL_HUFF_FAST_BITS) - 1u)];
    if (e & 0xff) { br_skip(b, (int)(e & 0xff)); return (int)(e >> 8); }
    // every code of up to UVOL_HUFF_FAST_BITS bits is in the fast table: the search starts behind it, on the bit-reversed window
    // (canonical codes are MSB-first, the stream is LSB-first), one compare per length
#if defined(__CUDA_ARCH__)
    const uint32_t rv = __brev(v);
#else
    uint32_t rv = 0; for (int i = 0; i < 32; i++) rv |= ((v >> i) & 1u) << (31 - i);
#endif
    // (all six candidate lengths are tested at once -- their table words are independent loads -- and the shortest match is kept:
    // a loop that stops at the first match would pay one shared-memory round trip per length, in series)
    uint32_t len = 0, idx = 0;
    const uint32_t maxl = T.maxl;
#pragma unroll
    for (uint32_t l = 16; l > UVOL_HUFF_FAST_BITS; l--) {
        const uint32_t fc = T.first_code[l], code = rv >> (32u - l), k = code - fc;
        if (l <= maxl && code >= fc && k < T.count[l]) { len = l; idx = T.first_idx[l] + k; }
    }
    if (len) { br_skip(b, (int)len); return sorted_pool[T.sorted_off + idx]; }
    br_skip(b, 16);
    return -1;
}

// Builds T from code sizes (serial part; the fast-table fill is split over `nlanes` callers).
UVOL_HD int huff_build_serial(HuffTable &T, const uint8_t *sizes, uint32_t total, uint16_t *sorted_pool, uint32_t sorted_off) {
    T.total = total; T.used = 0; T.maxl = 0; T.sorted_off = sorted_off;
    for (int l = 0; l < 17; l++) T.count[l] = 0;
    for (uint32_t i = 0; i < total; i++) { const uint32_t s = sizes[i]; if (s > 16) return UVOL_ERR_CORRUPT; if (s) { T.count[s]++; T.used++; if (s > T.maxl) T.maxl = s; } }
    uint32_t code = 0, idx = 0, fill[17];
    T.first_code[0] = 0; T.first_idx[0] = 0; fill[0] = 0;
    for (int l = 1; l <= 16; l++) { code = (code + (l > 1 ? T.count[l - 1] : 0)) << 1; T.first_code[l] = code; T.first_idx[l] = idx; fill[l] = idx; idx += T.count[l]; }
    for (uint32_t s = 0; s < total; s++) if (sizes[s]) sorted_pool[sorted_off + fill[sizes[s]]++] = (uint16_t)s;
    return UVOL_OK;
}
UVOL_HD void huff_fill_fast(HuffTable &T, const uint16_t *sorted_pool, uint32_t lane, uint32_t nlanes) {
    for (uint32_t i = lane; i < (1u << UVOL_HUFF_FAST_BITS); i += nlanes) T.fast[i] = 0;
}
UVOL_HD void huff_fill_fast2(HuffTable &T, const uint16_t *sorted_pool, uint32_t lane, uint32_t nlanes) {
    for (uint32_t l = 1; l <= UVOL_HUFF_FAST_BITS && l <= T.maxl; l++) {
        for (uint32_t k = lane; k < T.count[l]; k += nlanes) {
            const uint32_t c = T.first_code[l] + k; uint32_t rev = 0;
            for (uint32_t i = 0; i < l; i++) rev |= ((c >> i) & 1u) << (l - 1 - i);
            const uint32_t sym = sorted_pool[T.sorted_off + T.first_idx[l] + k];
            for (uint32_t hi = 0; hi < (1u << (UVOL_HUFF_FAST_BITS - l)); hi++) T.fast[rev | (hi << l)] = (sym << 8) | l;
        }
    }
}

// Reads one Huffman table header (code-length code + RLE'd lengths) into sizes[]; serial.
// `tmp` is a scratch HuffTable + 32-entry pool for the code-length code.
UVOL_HD int huff_read_sizes(BitRd &b, uint8_t *sizes, uint32_t max_total, uint32_t *out_total, HuffTable &tmp, uint16_t *tmp_pool) {
    const uint32_t total = br_get(b, 14);
    *out_total = total;
    if (total == 0) return UVOL_OK;
    if (total > max_total) return UVOL_ERR_CORRUPT;
    const uint32_t ncl = br_get(b, 5);
    if (ncl < 1 || ncl > 21) return UVOL_ERR_CORRUPT;
    const uint8_t order[21] = {17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16};
    uint8_t cls[21];
    for (int i = 0; i < 21; i++) cls[i] = 0;
    for (uint32_t i = 0; i < ncl; i++) cls[order[i]] = (uint8_t)br_get(b, 3);
    int rc = huff_build_serial(tmp, cls, 21, tmp_pool, 0); if (rc) return rc;
    huff_fill_fast(tmp, tmp_pool, 0, 1); huff_fill_fast2(tmp, tmp_pool, 0, 1);
    uint32_t n = 0;
    while (n < total) {
        const int c = huff_decode(tmp, tmp_pool, b);
        if (c < 0) return UVOL_ERR_CORRUPT;
        if (c <= 16) sizes[n++] = (uint8_t)c;
        else if (c == 17 || c == 18) { uint32_t rep = c == 17 ? br_get(b, 3) + 3 : br_get(b, 7) + 11; if (n + rep > total) return UVOL_ERR_CORRUPT; for (uint32_t k = 0; k < rep; k++) sizes[n++] = 0; }
        else { uint32_t rep = c == 19 ? br_get(b, 2) + 3 : br_get(b, 7) + 7; if (n == 0 || n + rep > total) return UVOL_ERR_CORRUPT; const uint8_t pv = sizes[n - 1]; for (uint32_t k = 0; k < rep; k++) sizes[n++] = pv; }
    }
    return UVOL_OK;
}

// ---------------------------------------------------------------------------------------------
// Slice symbol decode (B.3), serial per slice.  Emits per block: pred (2 bits), the delta-endpoint
// symbol (pred==3) and the selector index (non-CR blocks).  Endpoint indices are resolved later.
struct SliceTables { const HuffTable *epm, *dem, *sm, *rle; const uint16_t *pool; };
UVOL_HD int etc1s_slice_symbols(BitRd &b, const SliceTables &T, uint32_t bx, uint32_t by, uint32_t sel_count, uint32_t hist_size, int is_video,
                                uint8_t *rowp /*[bx]*/, uint16_t *hist /*[hist_size]*/, uint8_t *pred_out, uint16_t *delta_out, uint16_t *sel_out) {
    for (uint32_t i = 0; i < hist_size; i++) hist[i] = 0;
    uint32_t rover = hist_size / 2, rle_cnt = 0, prev_sym = 0, rep = 0, bits = 0;
    for (uint32_t y = 0; y < by; y++) {
        for (uint32_t x = 0; x < bx; x++) {
            if ((x & 1) == 0) {
                if ((y & 1) == 0) {
                    if (rep) { rep--; bits = prev_sym; }
                    else {
                        const int sy = huff_decode(*T.epm, T.pool, b); if (sy < 0) return UVOL_ERR_CORRUPT;
                        bits = (uint32_t)sy;
                        if (bits == 256) { rep = br_vlc(b, 4) + 3 - 1; bits = prev_sym; } else prev_sym = bits;
                    }
                    rowp[x] = (uint8_t)(bits >> 4);
                } else bits = rowp[x];
            }
            const uint32_t pred = bits & 3; bits >>= 2;
            const uint32_t bi = y * bx + x;
            pred_out[bi] = (uint8_t)pred;
            if (pred == 0 && x == 0) return UVOL_ERR_CORRUPT;
            if (pred == 1 && y == 0) return UVOL_ERR_CORRUPT;
            if (pred == 2 && !is_video && (x == 0 || y == 0)) return UVOL_ERR_CORRUPT;
            if (pred == 3) { const int d = huff_decode(*T.dem, T.pool, b); if (d < 0) return UVOL_ERR_CORRUPT; delta_out[bi] = (uint16_t)d; }
            if (!(is_video && pred == 2)) {
                uint32_t s;
                if (rle_cnt > 0) { rle_cnt--; s = hist[0]; }
                else {
                    const int sy = huff_decode(*T.sm, T.pool, b); if (sy < 0) return UVOL_ERR_CORRUPT;
                    s = (uint32_t)sy;
                    if (s == sel_count + hist_size) {
                        const int rr = huff_decode(*T.rle, T.pool, b); if (rr < 0) return UVOL_ERR_CORRUPT;
                        rle_cnt = (rr == 63) ? br_vlc(b, 7) + 3 : (uint32_t)rr + 3;
                        s = hist[0]; rle_cnt--;
                    } else if (s >= sel_count) {
                        const uint32_t i = s - sel_count; if (i >= hist_size) return UVOL_ERR_CORRUPT;
                        s = hist[i];
                        if (i) { const uint16_t t = hist[i / 2]; hist[i / 2] = hist[i]; hist[i] = t; }
                    } else { hist[rover++] = (uint16_t)s; if (rover == hist_size) rover = hist_size / 2; }
                }
                if (s >= sel_count) return UVOL_ERR_CORRUPT;
                sel_out[bi] = (uint16_t)s;
            }
        }
    }
    return UVOL_OK;
}

// ---------------------------------------------------------------------------------------------
// ETC1S block -> 16 RGBA texels (B.5).  ep = {r5,g5,b5,inten}, sel = 4 row bytes (2 bits/pixel).
// The intensity table {-b, -a, a, b} per table id: a = 2 5 9 13 18 24 33 47, b = 8 17 29 42 60 80 106 183, packed one byte per id
// (pure register arithmetic: an indexed local array would live in local memory and cost L1 traffic in the block kernels).
UVOL_HD uint32_t etc1s_color(uint32_t ep, int k) {
    const unsigned t8 = 8u * ((ep >> 24) & 7u);
    const int a = (int)((0x2F2118120D090502ull >> t8) & 255u), b = (int)((0xB76A503C2A1D1108ull >> t8) & 255u);
    const int d = k == 0 ? -b : (k == 1 ? -a : (k == 2 ? a : b));
    uint32_t out = 0;
    for (int c = 0; c < 3; c++) { const int c5 = (int)((ep >> (8 * c)) & 31); int v = ((c5 << 3) | (c5 >> 2)) + d; v = v < 0 ? 0 : (v > 255 ? 255 : v); out |= (uint32_t)v << (8 * c); }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Per-file global data (B.2): endpoint / selector codebooks and the four slice Huffman tables.
// Serial (a few thousand symbols per file); runs once per KTX2 file.
struct BasisGlobalsMem {
    uint32_t *endpoints;      // [ec]  r5 | g5<<8 | b5<<16 | inten<<24
    uint32_t *selectors;      // [sc]  4 row bytes, 2 bits per pixel
    HuffTable *tables;        // [0..3] endpoint_pred, delta_endpoint, selector, selector_rle ; [4..9] temporaries
    uint16_t *pool;           // sorted-symbol pool
    uint8_t *sizes;           // [32768] temporary code sizes
    uint32_t pool_cap;
};
UVOL_HD int basis_read_table(BitRd &b, HuffTable &T, BasisGlobalsMem &m, uint32_t &pool_used, uint32_t max_total) {
    uint32_t total = 0;
    HuffTable &tmp = m.tables[9]; uint16_t *tmp_pool = m.pool + m.pool_cap - 32;
    int rc = huff_read_sizes(b, m.sizes, max_total, &total, tmp, tmp_pool); if (rc) return rc;
    if (pool_used + total > m.pool_cap - 32) return UVOL_ERR_CORRUPT;
    rc = huff_build_serial(T, m.sizes, total, m.pool, pool_used); if (rc) return rc;
    huff_fill_fast(T, m.pool, 0, 1); huff_fill_fast2(T, m.pool, 0, 1);
    pool_used += T.used;
    return UVOL_OK;
}
UVOL_HD int basis_build_globals(const Ktx2File &f, const uint8_t *file, BasisGlobalsMem &m, uint32_t *out_hist_size) {
    uint32_t pool_used = 0; int rc;
    // slice tables first (they persist at the front of the pool)
    {
        BitRd b; br_init(b, file + f.tab_off, f.tab_len);
        if ((rc = basis_read_table(b, m.tables[0], m, pool_used, 32768))) return rc;
        if ((rc = basis_read_table(b, m.tables[1], m, pool_used, 32768))) return rc;
        if ((rc = basis_read_table(b, m.tables[2], m, pool_used, 32768))) return rc;
        if ((rc = basis_read_table(b, m.tables[3], m, pool_used, 32768))) return rc;
        const uint32_t hs = br_get(b, 13);
        if (hs == 0 || hs > 1024) return UVOL_ERR_UNSUPPORTED;
        *out_hist_size = hs;
        if ((b.consumed + 7) / 8 > f.tab_len) return UVOL_ERR_CORRUPT;
    }
    {   // endpoints
        BitRd b; br_init(b, file + f.ep_off, f.ep_len);
        for (int k = 0; k < 4; k++) if ((rc = basis_read_table(b, m.tables[4 + k], m, pool_used, 256))) return rc;
        const uint32_t gray = br_get(b, 1);
        uint32_t prev[3] = {16, 16, 16}, pint = 0;
        for (uint32_t i = 0; i < f.endpoint_count; i++) {
            int d = huff_decode(m.tables[7], m.pool, b); if (d < 0) return UVOL_ERR_CORRUPT;
            pint = (pint + (uint32_t)d) & 7;
            for (uint32_t c = 0; c < (gray ? 1u : 3u); c++) {
                const HuffTable &mt = prev[c] <= 9 ? m.tables[4] : (prev[c] <= 21 ? m.tables[5] : m.tables[6]);
                d = huff_decode(mt, m.pool, b); if (d < 0) return UVOL_ERR_CORRUPT;
                prev[c] = (prev[c] + (uint32_t)d) & 31;
            }
            if (gray) prev[1] = prev[2] = prev[0];
            m.endpoints[i] = prev[0] | (prev[1] << 8) | (prev[2] << 16) | (pint << 24);
        }
        if ((b.consumed + 7) / 8 > f.ep_len) return UVOL_ERR_CORRUPT;
    }
    {   // selectors
        BitRd b; br_init(b, file + f.sel_off, f.sel_len);
        const uint32_t glob = br_get(b, 1), hyb = br_get(b, 1), raw = br_get(b, 1);
        if (glob || hyb) return UVOL_ERR_UNSUPPORTED;
        if (raw) { for (uint32_t i = 0; i < f.selector_count; i++) { uint32_t v = 0; for (int j = 0; j < 4; j++) v |= br_get(b, 8) << (8 * j); m.selectors[i] = v; } }
        else {
            if ((rc = basis_read_table(b, m.tables[8], m, pool_used, 256))) return rc;
            uint32_t pb[4] = {0, 0, 0, 0};
            for (uint32_t i = 0; i < f.selector_count; i++) {
                uint32_t v = 0;
                for (int j = 0; j < 4; j++) {
                    if (i == 0) pb[j] = br_get(b, 8);
                    else { const int d = huff_decode(m.tables[8], m.pool, b); if (d < 0) return UVOL_ERR_CORRUPT; pb[j] ^= (uint32_t)d; }
                    v |= (pb[j] & 255u) << (8 * j);
                }
                m.selectors[i] = v;
            }
        }
        if ((b.consumed + 7) / 8 > f.sel_len) return UVOL_ERR_CORRUPT;
    }
    return UVOL_OK;
}

// One ETC1S block -> one ETC1 block (target format ETC1, src/lib/KTX2Loader.js:628-636): differential mode with a zero delta (both
// sub-blocks share the 5:5:5 base colour and the intensity table), flip bit set, selectors re-ordered to ETC1's pixel indices
// (column-major, value 0 1 2 3 = +a +b -a -b where the codebook stores 0 1 2 3 = -b -a +a +b).  Returned as the two 32-bit words a
// little-endian store writes: x = bytes 0-3 (R, G, B, table / diff / flip), y = bytes 4-7 (index MSBs, then LSBs, big-endian).
// ---- ETC2 RGBA target (UVOL_TEX_ETC2_RGBA): the 8-byte EAC alpha block that goes in front of the colour block.
// Replaces transcodeImage(..., ETC2, ...) for ETC1S sources with alpha (`etc2Supported`: [ETC1, ETC2] -> RGB_ETC2 / RGBA_ETC2_EAC,
// src/lib/KTX2Loader.js:619-627, the top-priority option for ETC1S).  The colour half is the ETC1 block as it is (differential mode with a
// zero delta is the same block in ETC2).  The alpha half: an ETC1S alpha block holds four values g + {-b, -a, a, b}; `map` (generated,
// tools/gen/gen_uastc_tables.py eac_map) gives per intensity table and set of used selectors the EAC {table, multiplier, base offset,
// index per selector} that reproduces them best.  Lossy by at most a few levels (bounds in tests/test_etc2.py); the reference's own
// table-driven conversion lives in the absent basis_transcoder WASM and cannot be bit-matched.
// EAC layout (big-endian 64 bits): base 8 | multiplier 4 | table 4 | 16 x 3-bit indices, pixel (x, y) at position x * 4 + y, first pixel in
// the top bits.  Returned as two little-endian words (bytes 0-3, 4-7).
struct EacWords { uint32_t x, y; };
UVOL_HD uint32_t eac_bswap(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
UVOL_HD EacWords eac_pack(uint32_t base, uint32_t mult, uint32_t table, unsigned long long idx48) {
    const unsigned long long be = ((unsigned long long)base << 56) | ((unsigned long long)mult << 52) | ((unsigned long long)table << 48) | idx48;
    EacWords o; o.x = eac_bswap((uint32_t)(be >> 32)); o.y = eac_bswap((uint32_t)be);
    return o;
}
UVOL_HD EacWords eac_opaque() { return eac_pack(255u, 1u, 13u, 0x924924924924ull); }      // table 13, index 4 (modifier 0) everywhere: 255
UVOL_HD EacWords etc1s_alpha_to_eac(uint32_t aep, uint32_t asel, const uint32_t *map) {
    uint32_t used = 0;
    for (int i = 0; i < 16; i++) used |= 1u << ((asel >> (2 * i)) & 3u);
    const uint32_t g5 = (aep >> 8) & 31u, g = (g5 << 3) | (g5 >> 2), e = map[((aep >> 24) & 7u) * 16u + used];
    int base = (int)g + (int)((e >> 8) & 255u) - 128; base = base < 0 ? 0 : (base > 255 ? 255 : base);
    unsigned long long idx = 0;
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
        const uint32_t q = (asel >> (8 * y + 2 * x)) & 3u, j = (e >> (16u + 3u * q)) & 7u;
        idx |= (unsigned long long)j << (45 - 3 * (x * 4 + y));
    }
    return eac_pack((uint32_t)base, (e >> 4) & 15u, e & 15u, idx);
}

// ---- BC1 / BC3 targets (UVOL_TEX_BC1 / UVOL_TEX_BC3): `dxtSupported`, transcoderFormat [BC1, BC3] -> RGB_S3TC_DXT1 / RGBA_S3TC_DXT5
// (src/lib/KTX2Loader.js:610-618), the reference's fallback on desktop GPUs without BPTC.  ETC1S sources: the block's darkest and
// brightest SELECTED colours become the two RGB565 endpoints, every selector takes the nearest of the four BC1 colours; the BC3 alpha
// block (BC4: two 8-bit endpoints, eight interpolants, 3-bit indices) is built the same way from the alpha slice.  Lossy (565
// endpoints, thirds instead of ETC1S's intensity steps); decoded by Pillow's DXT1 / DXT5 decoder in tests/test_dxt.py, PSNR bounds
// there.  Index layouts coincide with ETC1S's selector layout (pixel (x, y) at field 4y + x).
UVOL_HD uint32_t bc1_q(uint32_t v, uint32_t bits) {                 // nearest code under MSB replication
    const uint32_t maxq = (1u << bits) - 1u; uint32_t q = (v * maxq + 127u) / 255u, best = q, beste = 0xffffu;
    for (uint32_t c = q ? q - 1u : 0u; c <= (q < maxq ? q + 1u : maxq); c++) {
        const uint32_t e8 = bits == 5u ? ((c << 3) | (c >> 2)) : ((c << 2) | (c >> 4)), e = e8 > v ? e8 - v : v - e8;
        if (e < beste) { beste = e; best = c; }
    }
    return best;
}
UVOL_HD uint32_t bc1_565(uint32_t rgb) { return (bc1_q(rgb & 255u, 5) << 11) | (bc1_q((rgb >> 8) & 255u, 6) << 5) | bc1_q((rgb >> 16) & 255u, 5); }
UVOL_HD uint32_t bc1_rgb(uint32_t c) {                              // RGB565 -> packed RGB8
    const uint32_t r = c >> 11, g = (c >> 5) & 63u, b = c & 31u;
    return ((r << 3) | (r >> 2)) | (((g << 2) | (g >> 4)) << 8) | (((b << 3) | (b >> 2)) << 16);
}
UVOL_HD uint32_t bc1_dist(uint32_t a, uint32_t b) {
    uint32_t d = 0;
    for (int c = 0; c < 3; c++) { const uint32_t x = (a >> (8 * c)) & 255u, y = (b >> (8 * c)) & 255u; d += x > y ? x - y : y - x; }
    return d;
}
UVOL_HD uint32_t bc1_mix(uint32_t a, uint32_t b) {                  // (2a + b) / 3 per channel
    uint32_t o = 0;
    for (int c = 0; c < 3; c++) o |= ((2u * ((a >> (8 * c)) & 255u) + ((b >> (8 * c)) & 255u)) / 3u) << (8 * c);
    return o;
}
struct Bc1Words { uint32_t x, y; };
UVOL_HD Bc1Words etc1s_to_bc1(uint32_t ep, uint32_t sel) {
    uint32_t col[4], used = 0;
    for (int k = 0; k < 4; k++) col[k] = etc1s_color(ep, k);
    for (int i = 0; i < 16; i++) used |= 1u << ((sel >> (2 * i)) & 3u);
    const uint32_t smin = used & 1u ? 0u : (used & 2u ? 1u : (used & 4u ? 2u : 3u)), smax = used & 8u ? 3u : (used & 4u ? 2u : (used & 2u ? 1u : 0u));
    const uint32_t ca = bc1_565(col[smin]), cb = bc1_565(col[smax]);
    Bc1Words o;
    if (ca == cb) { o.x = ca | (cb << 16); o.y = 0; return o; }     // c0 == c1: index 0 everywhere is that colour
    const uint32_t c0 = ca > cb ? ca : cb, c1 = ca > cb ? cb : ca;   // c0 > c1: the four-colour mode
    uint32_t P[4]; P[0] = bc1_rgb(c0); P[1] = bc1_rgb(c1); P[2] = bc1_mix(P[0], P[1]); P[3] = bc1_mix(P[1], P[0]);
    uint32_t map = 0;
    for (uint32_t s = 0; s < 4; s++) {
        uint32_t bestj = 0, beste = 0xffffffffu;
        for (uint32_t j = 0; j < 4; j++) { const uint32_t e = bc1_dist(P[j], col[s]); if (e < beste) { beste = e; bestj = j; } }
        map |= bestj << (2u * s);
    }
    uint32_t idx = 0;
    for (int i = 0; i < 16; i++) idx |= ((map >> (2u * ((sel >> (2 * i)) & 3u))) & 3u) << (2 * i);
    o.x = c0 | (c1 << 16); o.y = idx;
    return o;
}
// BC4 alpha block of BC3 from an ETC1S alpha-slice block (alpha = G of its four colours); bc4_opaque: 255 everywhere
UVOL_HD Bc1Words bc4_opaque() { Bc1Words o; o.x = 0xffffu; o.y = 0; return o; }
UVOL_HD Bc1Words etc1s_alpha_to_bc4(uint32_t aep, uint32_t asel) {
    uint32_t av[4], used = 0;
    for (int k = 0; k < 4; k++) av[k] = (etc1s_color(aep, k) >> 8) & 255u;
    for (int i = 0; i < 16; i++) used |= 1u << ((asel >> (2 * i)) & 3u);
    const uint32_t smin = used & 1u ? 0u : (used & 2u ? 1u : (used & 4u ? 2u : 3u)), smax = used & 8u ? 3u : (used & 4u ? 2u : (used & 2u ? 1u : 0u));
    const uint32_t a0 = av[smax], a1 = av[smin];
    Bc1Words o;
    if (a0 <= a1) { o.x = a0 | (a0 << 8); o.y = 0; return o; }      // one value (a0 == a1 selects the six-value mode: index 0 is a0)
    uint32_t map = 0;
    for (uint32_t s = 0; s < 4; s++) {
        uint32_t bestj = 0, beste = 0xffffffffu;
        for (uint32_t j = 0; j < 8; j++) {
            const uint32_t v = j == 0 ? a0 : (j == 1 ? a1 : ((8u - j) * a0 + (j - 1u) * a1) / 7u), e = v > av[s] ? v - av[s] : av[s] - v;
            if (e < beste) { beste = e; bestj = j; }
        }
        map |= bestj << (3u * s);
    }
    unsigned long long idx = 0;
    for (int i = 0; i < 16; i++) idx |= (unsigned long long)((map >> (3u * ((asel >> (2 * i)) & 3u))) & 7u) << (3 * i);
    o.x = a0 | (a1 << 8) | ((uint32_t)(idx & 0xffffu) << 16); o.y = (uint32_t)(idx >> 16);
    return o;
}

struct Etc1Words { uint32_t x, y; };
UVOL_HD Etc1Words etc1s_to_etc1(uint32_t ep, uint32_t sel) {
    const uint32_t r5 = ep & 31u, g5 = (ep >> 8) & 31u, b5 = (ep >> 16) & 31u, inten = (ep >> 24) & 7u;
    Etc1Words o;
    o.x = (r5 << 3) | (g5 << 11) | (b5 << 19) | (((inten << 5) | (inten << 2) | 3u) << 24);
    uint32_t msb = 0, lsb = 0;
    for (int y = 0; y < 4; y++) {
        const uint32_t rb = (sel >> (8 * y)) & 255u;
        for (int x = 0; x < 4; x++) {
            const uint32_t q = (rb >> (2 * x)) & 3u, e = (0x4Bu >> (2 * q)) & 3u;          // 0 1 2 3 -> 3 2 0 1
            const int p = x * 4 + y;
            msb |= (e >> 1) << p; lsb |= (e & 1u) << p;
        }
    }
    o.y = (msb >> 8) | ((msb & 255u) << 8) | ((lsb >> 8) << 16) | ((lsb & 255u) << 24);
    return o;
}

// One ETC1S block -> RGBA32 rows.  rows[r] receives 4 packed RGBA texels of pixel row r.
UVOL_HD void etc1s_block_rows(uint32_t ep, uint32_t sel, uint32_t rows[4][4]) {
    const uint32_t c0 = etc1s_color(ep, 0) | 0xff000000u, c1 = etc1s_color(ep, 1) | 0xff000000u, c2 = etc1s_color(ep, 2) | 0xff000000u, c3 = etc1s_color(ep, 3) | 0xff000000u;
    for (int y = 0; y < 4; y++) {
        const uint32_t rb = (sel >> (8 * y)) & 255u;
        for (int x = 0; x < 4; x++) { const uint32_t q = (rb >> (2 * x)) & 3u; rows[y][x] = (q & 2u) ? ((q & 1u) ? c3 : c2) : ((q & 1u) ? c1 : c0); }      // selects, not an indexed array
    }
}
