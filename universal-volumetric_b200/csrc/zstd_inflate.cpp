// zstd_inflate.cpp -- Zstandard frame decoder (RFC 8878) for KTX2 supercompressionScheme 2.
//
// The reference inflates Zstd-supercompressed KTX2 levels on the CPU before transcoding (src/lib/zstddec.module.js, a WASM
// build of the zstd decoder; src/lib/KTX2Loader.js:803-817 for plain textures, inside the Basis module for UASTC levels --
// `basisu -uastc -ktx2` writes Zstd levels by default).  Here it is host code in front of the UASTC block kernel: one level
// = one frame, inflated by a pool of host threads straight into the pinned staging blob that is uploaded to the GPU.
// Written from the format specification; supports everything a conforming encoder emits except dictionaries:
// raw / RLE / compressed blocks, raw / RLE / Huffman (1 or 4 streams, direct or FSE-coded weights) / treeless literals,
// predefined / RLE / FSE / repeat sequence tables, repeat offsets, skippable and concatenated frames.  The content checksum
// is skipped, not verified.  Checked byte for byte against libzstd 1.5.5 in tests/test_zstd.py (test-only use of libzstd).
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <atomic>
#include "uvol_internal.h"

namespace {

// Feature counters (test aid: which parts of the format the vectors exercised; not synchronised).
// 0 raw blocks, 1 RLE blocks, 2 compressed blocks, 3 raw literals, 4 RLE literals, 5 Huffman literals, 6 treeless literals, 7 four-stream
// literals, 8 direct weights, 9 FSE-coded weights, 10 predefined tables, 11 RLE tables, 12 FSE tables, 13 repeat tables, 14 repeat offsets, 15 frames
std::atomic<uint64_t> g_feat[16];          // test aid (feature coverage); relaxed atomics: the staging threads inflate concurrently

struct FseEntry { uint8_t sym, nbits; uint16_t base; };          // next state = base + read(nbits)
struct FseTable { FseEntry e[512]; int log; };                    // accuracy log <= 9
struct HufTable { uint16_t e[2048]; int maxbits; };               // (sym << 8) | nbits, indexed by the top maxbits (<= 11) bits
struct SeqTables { FseTable ll, of, ml; bool have_ll = false, have_of = false, have_ml = false; };

inline int highbit(uint32_t v) { return 31 - __builtin_clz(v); }

// ---- forward bit reader (FSE table descriptions)
struct FwdBits { const uint8_t *p; size_t n; uint64_t pos; };     // pos in bits
inline uint32_t fwd_peek(const FwdBits &b, int nb) {
    uint64_t v = 0; const size_t byte = b.pos >> 3;
    for (int i = 0; i < 5; i++) if (byte + i < b.n) v |= (uint64_t)b.p[byte + i] << (8 * i);
    return (uint32_t)((v >> (b.pos & 7)) & ((1ull << nb) - 1));
}

// ---- backward bit reader (Huffman streams, FSE weights, sequences): the stream is read from its last byte towards its first;
// `bits` = number of unread bits, bit k of the stream = bit (k & 7) of byte k >> 3.  Reading past the first byte yields zeros.
struct BackBits { const uint8_t *p; size_t n; int64_t bits; };
inline bool back_init(BackBits &b, const uint8_t *p, size_t n) {
    if (n == 0 || p[n - 1] == 0) return false;
    b.p = p; b.n = n; b.bits = (int64_t)n * 8 - (8 - highbit(p[n - 1]));        // drop the padding and the marker bit
    return true;
}
inline uint64_t back_read(BackBits &b, int nb) {                  // nb <= 32
    if (nb == 0) return 0;
    const int64_t lo = b.bits - nb;                               // index of the lowest bit wanted
    b.bits = lo;
    uint64_t v = 0;
    if (lo >= 0) {
        const size_t byte = (size_t)(lo >> 3);
        if (byte + 8 <= b.n) memcpy(&v, b.p + byte, 8); else memcpy(&v, b.p + byte, b.n - byte);
        return (v >> (lo & 7)) & ((1ull << nb) - 1);
    }
    const int64_t have = nb + lo;                                 // bits that exist, the rest reads as zero
    if (have <= 0) return 0;
    memcpy(&v, b.p, b.n < 8 ? b.n : 8);
    return (v & ((1ull << have) - 1)) << (nb - have);
}

// ---- FSE
bool fse_build(FseTable &t, const int16_t *norm, int nsym, int log) {
    if (log > 9 || log < 1) return false;
    const int size = 1 << log; t.log = log;
    uint16_t next[256]; int high = size - 1;
    for (int s = 0; s < nsym; s++) {
        if (norm[s] == -1) { t.e[high--].sym = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1; int pos = 0;
    for (int s = 0; s < nsym; s++) {
        for (int i = 0; i < norm[s]; i++) { t.e[pos].sym = (uint8_t)s; do { pos = (pos + step) & mask; } while (pos > high); }
    }
    if (pos != 0) return false;
    for (int i = 0; i < size; i++) {
        const uint8_t s = t.e[i].sym; const uint32_t ns = next[s]++;
        const int nb = log - highbit(ns);
        t.e[i].nbits = (uint8_t)nb; t.e[i].base = (uint16_t)((ns << nb) - size);
    }
    return true;
}
// Reads a normalised distribution (RFC 8878 4.1.1).  Returns bytes consumed, 0 on error.
size_t fse_read_norm(const uint8_t *p, size_t n, int16_t *norm, int max_sym, int max_log, int *nsym_out, int *log_out) {
    if (n < 1) return 0;
    FwdBits b{p, n, 0};
    const int log = (int)fwd_peek(b, 4) + 5; b.pos += 4;
    if (log > max_log) return 0;
    int remaining = (1 << log) + 1, threshold = 1 << log, nbits = log + 1, sym = 0; bool prev0 = false;
    while (remaining > 1 && sym <= max_sym) {
        if (prev0) {
            int rep;
            do { rep = (int)fwd_peek(b, 2); b.pos += 2; for (int i = 0; i < rep && sym <= max_sym; i++) norm[sym++] = 0; } while (rep == 3);
            prev0 = false;
            if (sym > max_sym) break;
            continue;
        }
        const int max = (2 * threshold - 1) - remaining;
        int count;
        const uint32_t v = fwd_peek(b, nbits);
        if ((int)(v & (uint32_t)(threshold - 1)) < max) { count = (int)(v & (uint32_t)(threshold - 1)); b.pos += (uint64_t)(nbits - 1); }
        else { count = (int)(v & (uint32_t)(2 * threshold - 1)); if (count >= threshold) count -= max; b.pos += (uint64_t)nbits; }
        count--;                                                   // -1 = "less than one"
        remaining -= count < 0 ? -count : count;
        norm[sym++] = (int16_t)count;
        prev0 = count == 0;
        while (remaining < threshold) { nbits--; threshold >>= 1; }
        if ((b.pos + 7) / 8 > n + 4) return 0;
    }
    if (remaining != 1 || sym > max_sym + 1) return 0;
    *nsym_out = sym; *log_out = log;
    const size_t used = (size_t)((b.pos + 7) / 8);
    return used <= n ? used : 0;
}

// ---- Huffman
// weights[0..nsym): builds the decoding table.  The last weight is implied (total must complete a power of two).
bool huf_build(HufTable &t, uint8_t *w, int nsym) {
    uint32_t total = 0;
    for (int i = 0; i < nsym; i++) { if (w[i] > 11) return false; total += w[i] ? (1u << (w[i] - 1)) : 0; }
    if (total == 0) return false;
    const int maxbits = highbit(total) + 1;
    if (maxbits > 11) return false;
    const uint32_t left = (1u << maxbits) - total;
    if (left == 0 || (left & (left - 1))) return false;           // must be a power of two
    w[nsym] = (uint8_t)(highbit(left) + 1); nsym++;
    if (nsym > 256) return false;
    t.maxbits = maxbits;
    uint32_t rank_count[13] = {0}, rank_start[13];
    for (int i = 0; i < nsym; i++) rank_count[w[i]]++;
    uint32_t pos = 0;
    for (int r = 1; r <= maxbits; r++) { rank_start[r] = pos; pos += rank_count[r] << (r - 1); }
    if (pos != (1u << maxbits)) return false;
    for (int s = 0; s < nsym; s++) {
        const int r = w[s]; if (!r) continue;
        const uint32_t len = 1u << (r - 1); const uint16_t v = (uint16_t)((s << 8) | (maxbits + 1 - r));
        for (uint32_t k = 0; k < len; k++) t.e[rank_start[r] + k] = v;
        rank_start[r] += len;
    }
    return true;
}
// Huffman tree description (4.2.1).  Returns bytes consumed, 0 on error.
size_t huf_read_table(const uint8_t *p, size_t n, HufTable &t) {
    if (n < 1) return 0;
    uint8_t w[257]; int nsym = 0; size_t used;
    const int hb = p[0];
    g_feat[hb >= 128 ? 8 : 9].fetch_add(1, std::memory_order_relaxed);
    if (hb >= 128) {                                               // direct 4-bit weights
        nsym = hb - 127; used = 1 + (size_t)(nsym + 1) / 2;
        if (used > n) return 0;
        for (int i = 0; i < nsym; i++) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
    } else {                                                       // FSE-compressed weights, two interleaved states
        used = 1 + (size_t)hb;
        if (hb == 0 || used > n) return 0;
        int16_t norm[16]; int ns, log;
        const size_t hdr = fse_read_norm(p + 1, (size_t)hb, norm, 12, 6, &ns, &log);
        if (!hdr || hdr >= (size_t)hb) return 0;
        FseTable ft; if (!fse_build(ft, norm, ns, log)) return 0;
        BackBits b; if (!back_init(b, p + 1 + hdr, (size_t)hb - hdr)) return 0;
        uint32_t s1 = (uint32_t)back_read(b, log), s2 = (uint32_t)back_read(b, log);
        for (;;) {
            if (nsym >= 255) return 0;
            w[nsym++] = ft.e[s1].sym;
            if (b.bits < ft.e[s1].nbits) { if (nsym >= 255) return 0; w[nsym++] = ft.e[s2].sym; break; }
            s1 = ft.e[s1].base + (uint32_t)back_read(b, ft.e[s1].nbits);
            if (nsym >= 255) return 0;
            w[nsym++] = ft.e[s2].sym;
            if (b.bits < ft.e[s2].nbits) { if (nsym >= 255) return 0; w[nsym++] = ft.e[s1].sym; break; }
            s2 = ft.e[s2].base + (uint32_t)back_read(b, ft.e[s2].nbits);
        }
    }
    if (!huf_build(t, w, nsym)) return 0;
    return used;
}
// Fast backward reader for the Huffman streams: a 64-bit container loaded from the byte pointer, `consumed` counted from its top.
struct BackFast {
    const uint8_t *start, *ptr; uint64_t c; unsigned consumed;
    bool init(const uint8_t *p, size_t n) {
        if (n == 0 || p[n - 1] == 0) return false;
        start = p;
        if (n >= 8) { ptr = p + n - 8; memcpy(&c, ptr, 8); consumed = 8u - (unsigned)highbit(p[n - 1]); }
        else { ptr = p; c = 0; for (size_t i = 0; i < n; i++) c |= (uint64_t)p[i] << (8 * i); consumed = 8u - (unsigned)highbit(p[n - 1]) + (unsigned)(8 - n) * 8u; }
        return true;
    }
    inline void reload() {
        if (ptr - start >= 8) { ptr -= consumed >> 3; consumed &= 7u; memcpy(&c, ptr, 8); }
        else if (ptr != start) { size_t nb = consumed >> 3; if (nb > (size_t)(ptr - start)) nb = (size_t)(ptr - start); ptr -= nb; consumed -= (unsigned)(8 * nb); memcpy(&c, ptr, 8); }
    }
    inline uint64_t read(unsigned nb) {                            // nb <= 32; call reload() so that the bits are in the container; past the first byte: zeros
        if (consumed >= 64) { consumed += nb; return 0; }
        const uint64_t v = ((c << consumed) >> 1) >> (63u - nb);
        consumed += nb;
        return v;
    }
    inline bool roomy() const { return ptr - start >= 8; }         // a reload brings at least 57 unread bits
    inline bool done() const { return ptr == start && consumed == 64; }
};
#define HUF_SYM(B, O) do { const uint16_t e_ = t.e[((B).c << (B).consumed) >> sh]; *(O)++ = (uint8_t)(e_ >> 8); (B).consumed += e_ & 255u; } while (0)
inline bool huf_tail(const HufTable &t, BackFast &b, uint8_t *o, uint8_t *end) {
    const unsigned sh = 64u - (unsigned)t.maxbits;
    while (o < end) {
        b.reload();
        if (b.consumed >= 64) return false;                        // symbols left but no bits
        HUF_SYM(b, o);
        if (b.consumed > 64) return false;
    }
    b.reload();
    return b.done();
}
bool huf_decode_stream(const HufTable &t, const uint8_t *p, size_t n, uint8_t *out, size_t count) {
    BackFast b; if (!b.init(p, n)) return false;
    const unsigned sh = 64u - (unsigned)t.maxbits;
    uint8_t *o = out, *end = out + count;
    while (end - o >= 5 && b.roomy()) { b.reload(); HUF_SYM(b, o); HUF_SYM(b, o); HUF_SYM(b, o); HUF_SYM(b, o); HUF_SYM(b, o); }      // 5 x 11 bits <= 57
    return huf_tail(t, b, o, end);
}
// Four streams in lock step: four independent dependency chains keep the core busy.
bool huf_decode_4(const HufTable &t, const uint8_t *const p[4], const size_t n[4], uint8_t *const out[4], const size_t count[4]) {
    BackFast b0, b1, b2, b3;
    if (!b0.init(p[0], n[0]) || !b1.init(p[1], n[1]) || !b2.init(p[2], n[2]) || !b3.init(p[3], n[3])) return false;
    const unsigned sh = 64u - (unsigned)t.maxbits;
    uint8_t *o0 = out[0], *o1 = out[1], *o2 = out[2], *o3 = out[3];
    uint8_t *e0 = o0 + count[0], *e1 = o1 + count[1], *e2 = o2 + count[2], *e3 = o3 + count[3];
    while (e0 - o0 >= 5 && e1 - o1 >= 5 && e2 - o2 >= 5 && e3 - o3 >= 5 && b0.roomy() && b1.roomy() && b2.roomy() && b3.roomy()) {
        b0.reload(); b1.reload(); b2.reload(); b3.reload();
        for (int k = 0; k < 5; k++) { HUF_SYM(b0, o0); HUF_SYM(b1, o1); HUF_SYM(b2, o2); HUF_SYM(b3, o3); }
    }
    return huf_tail(t, b0, o0, e0) && huf_tail(t, b1, o1, e1) && huf_tail(t, b2, o2, e2) && huf_tail(t, b3, o3, e3);
}
#undef HUF_SYM

// ---- sequences
const int16_t LL_DEF[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
const int16_t ML_DEF[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
const int16_t OF_DEF[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
const uint32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
const uint8_t LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
const uint32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
const uint8_t ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

// One of the three sequence tables according to its 2-bit mode.  Returns bytes consumed (may be 0), (size_t)-1 on error.
size_t seq_table(int mode, const uint8_t *p, size_t n, FseTable &t, bool &have, const int16_t *def, int def_n, int def_log, int max_sym, int max_log) {
    g_feat[10 + mode].fetch_add(1, std::memory_order_relaxed);
    if (mode == 0) { if (!fse_build(t, def, def_n, def_log)) return (size_t)-1; have = true; return 0; }
    if (mode == 1) { if (n < 1 || p[0] > max_sym) return (size_t)-1; t.log = 0; t.e[0].sym = p[0]; t.e[0].nbits = 0; t.e[0].base = 0; have = true; return 1; }
    if (mode == 2) {
        int16_t norm[64]; int ns, log;
        const size_t used = fse_read_norm(p, n, norm, max_sym, max_log, &ns, &log);
        if (!used || !fse_build(t, norm, ns, log)) return (size_t)-1;
        have = true; return used;
    }
    return have ? 0 : (size_t)-1;                                  // repeat
}

struct FrameState { HufTable huf; bool have_huf = false; SeqTables seq; uint64_t rep[3] = {1, 4, 8}; };

// One compressed block: literals + sequences -> dst[*dpos ...].  false on any format violation.
bool block_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *dpos, FrameState &fs, uint8_t *lit_buf) {
    if (n < 1) return false;
    // ---- literals section header
    const int ltype = src[0] & 3, sf = (src[0] >> 2) & 3;
    size_t regen, csize = 0, hdr; int streams = 1;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { regen = src[0] >> 3; hdr = 1; }
        else if (sf == 1) { if (n < 2) return false; regen = (src[0] >> 4) | ((size_t)src[1] << 4); hdr = 2; }
        else { if (n < 3) return false; regen = (src[0] >> 4) | ((size_t)src[1] << 4) | ((size_t)src[2] << 12); hdr = 3; }
    } else {
        if (sf < 2) { if (n < 3) return false; const uint32_t v = src[0] | (src[1] << 8) | ((uint32_t)src[2] << 16); regen = (v >> 4) & 1023; csize = (v >> 14) & 1023; hdr = 3; streams = sf == 0 ? 1 : 4; }
        else if (sf == 2) { if (n < 4) return false; const uint32_t v = src[0] | (src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24); regen = (v >> 4) & 16383; csize = v >> 18; hdr = 4; streams = 4; }
        else { if (n < 5) return false; const uint64_t v = src[0] | (src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint64_t)src[3] << 24) | ((uint64_t)src[4] << 32); regen = (size_t)((v >> 4) & 262143); csize = (size_t)(v >> 22); hdr = 5; streams = 4; }
    }
    if (regen > (128u << 10)) return false;
    g_feat[3 + ltype].fetch_add(1, std::memory_order_relaxed); if (streams == 4) g_feat[7].fetch_add(1, std::memory_order_relaxed);
    const uint8_t *lit; size_t p = hdr;
    if (ltype == 0) { if (p + regen > n) return false; lit = src + p; p += regen; }
    else if (ltype == 1) { if (p + 1 > n) return false; memset(lit_buf, src[p], regen); lit = lit_buf; p += 1; }
    else {
        if (p + csize > n) return false;
        const uint8_t *c = src + p; size_t cn = csize;
        if (ltype == 2) { const size_t used = huf_read_table(c, cn, fs.huf); if (!used) return false; fs.have_huf = true; c += used; cn -= used; }
        else if (!fs.have_huf) return false;
        if (streams == 1) { if (!huf_decode_stream(fs.huf, c, cn, lit_buf, regen)) return false; }
        else {
            if (cn < 6) return false;
            const size_t s1 = c[0] | (c[1] << 8), s2 = c[2] | (c[3] << 8), s3 = c[4] | (c[5] << 8);
            if (6 + s1 + s2 + s3 > cn) return false;
            const size_t s4 = cn - 6 - s1 - s2 - s3, q = (regen + 3) / 4;
            if (3 * q > regen) return false;
            const uint8_t *b = c + 6;
            const uint8_t *const sp[4] = {b, b + s1, b + s1 + s2, b + s1 + s2 + s3}; const size_t sn[4] = {s1, s2, s3, s4};
            uint8_t *const so[4] = {lit_buf, lit_buf + q, lit_buf + 2 * q, lit_buf + 3 * q}; const size_t sc[4] = {q, q, q, regen - 3 * q};
            if (!huf_decode_4(fs.huf, sp, sn, so, sc)) return false;
        }
        lit = lit_buf; p += csize;
    }
    // ---- sequences section
    if (p >= n) {                                                  // no sequences section at all is only legal as "0 sequences"
        if (p > n) return false;
    }
    size_t nseq = 0;
    if (p < n) {
        const int b0 = src[p];
        if (b0 < 128) { nseq = (size_t)b0; p += 1; }
        else if (b0 < 255) { if (p + 2 > n) return false; nseq = ((size_t)(b0 - 128) << 8) + src[p + 1]; p += 2; }
        else { if (p + 3 > n) return false; nseq = (size_t)src[p + 1] + ((size_t)src[p + 2] << 8) + 0x7F00; p += 3; }
    }
    size_t d = *dpos, lpos = 0;
    if (nseq) {
        if (p + 1 > n) return false;
        const int modes = src[p++];
        if (modes & 3) return false;
        size_t u;
        if ((u = seq_table((modes >> 6) & 3, src + p, n - p, fs.seq.ll, fs.seq.have_ll, LL_DEF, 36, 6, 35, 9)) == (size_t)-1) return false;
        p += u;
        if ((u = seq_table((modes >> 4) & 3, src + p, n - p, fs.seq.of, fs.seq.have_of, OF_DEF, 29, 5, 31, 8)) == (size_t)-1) return false;
        p += u;
        if ((u = seq_table((modes >> 2) & 3, src + p, n - p, fs.seq.ml, fs.seq.have_ml, ML_DEF, 53, 6, 52, 9)) == (size_t)-1) return false;
        p += u;
        BackFast b; if (!b.init(src + p, n - p)) return false;
        const FseTable &LL = fs.seq.ll, &OF = fs.seq.of, &ML = fs.seq.ml;
        uint32_t sl = (uint32_t)b.read((unsigned)LL.log), so = (uint32_t)b.read((unsigned)OF.log), sm = (uint32_t)b.read((unsigned)ML.log);   // <= 26 bits
        // 16-byte copies may overshoot: the literal source is readable up to lit_lim, the destination up to cap
        const uint8_t *const lit_lim = lit == lit_buf ? lit_buf + (128u << 10) + 32 : src + n;
        for (size_t i = 0; i < nseq; i++) {
            const FseEntry el = LL.e[sl], eo = OF.e[so], em = ML.e[sm];
            const unsigned lc = el.sym, oc = eo.sym, mc = em.sym;
            if (lc > 35 || mc > 52 || oc > 31) return false;
            b.reload();
            const uint64_t ov = (1ull << oc) + b.read(oc);                       // <= 31 bits
            if (oc + ML_BITS[mc] + LL_BITS[lc] > 56) b.reload();
            const size_t mlen = ML_BASE[mc] + (size_t)b.read(ML_BITS[mc]);       // <= 16 + 16 bits
            const size_t llen = LL_BASE[lc] + (size_t)b.read(LL_BITS[lc]);
            uint64_t offset;
            if (ov > 3) { offset = ov - 3; fs.rep[2] = fs.rep[1]; fs.rep[1] = fs.rep[0]; fs.rep[0] = offset; }
            else {
                g_feat[14].fetch_add(1, std::memory_order_relaxed);
                const uint64_t idx = ov - 1 + (llen == 0 ? 1 : 0);          // 0..3
                if (idx == 0) offset = fs.rep[0];
                else {
                    offset = idx < 3 ? fs.rep[idx] : fs.rep[0] - 1;
                    if (offset == 0) return false;
                    if (idx > 1) fs.rep[2] = fs.rep[1];
                    fs.rep[1] = fs.rep[0]; fs.rep[0] = offset;
                }
            }
            if (i + 1 < nseq) {                                    // state updates: literals length, match length, offset (<= 26 bits)
                b.reload();
                sl = el.base + (uint32_t)b.read(el.nbits);
                sm = em.base + (uint32_t)b.read(em.nbits);
                so = eo.base + (uint32_t)b.read(eo.nbits);
            }
            if (lpos + llen > regen || d + llen + mlen > cap || offset > d + llen) return false;
            uint8_t *o = dst + d; const uint8_t *ls = lit + lpos;
            if (llen <= 16 && ls + 16 <= lit_lim && d + 16 <= cap) memcpy(o, ls, 16); else memcpy(o, ls, llen);
            d += llen; lpos += llen; o += llen;
            const uint8_t *m = o - offset;
            if (offset >= 16 && d + mlen + 16 <= cap) { for (size_t k = 0; k < mlen; k += 16) memcpy(o + k, m + k, 16); }
            else if (offset >= mlen) memcpy(o, m, mlen);
            else for (size_t k = 0; k < mlen; k++) o[k] = m[k];
            d += mlen;
        }
        b.reload();
        if (!b.done()) return false;
    }
    if (d + (regen - lpos) > cap) return false;
    memcpy(dst + d, lit + lpos, regen - lpos); d += regen - lpos;
    *dpos = d;
    return true;
}

}  // namespace

extern "C" void uvol_zstd_feature_counts(uint64_t *out16, int reset) { for (int i = 0; i < 16; i++) { out16[i] = g_feat[i].load(std::memory_order_relaxed); if (reset) g_feat[i].store(0, std::memory_order_relaxed); } }

// Inflates every frame in src[0..n) into dst (capacity cap).  Returns UVOL_OK and the byte count, or a negative status.
extern "C" int uvol_zstd_inflate(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len) {
    if (!src || (!dst && cap) || !out_len) return UVOL_ERR_ARG;
    size_t p = 0, d = 0;
    uint8_t *lit_buf = new uint8_t[(128u << 10) + 32];
    int rc = UVOL_OK;
    while (p < n && rc == UVOL_OK) {
        if (n - p < 4) { rc = UVOL_ERR_TRUNCATED; break; }
        const uint32_t magic = src[p] | (src[p + 1] << 8) | ((uint32_t)src[p + 2] << 16) | ((uint32_t)src[p + 3] << 24);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {                 // skippable frame
            if (n - p < 8) { rc = UVOL_ERR_TRUNCATED; break; }
            const uint32_t sz = src[p + 4] | (src[p + 5] << 8) | ((uint32_t)src[p + 6] << 16) | ((uint32_t)src[p + 7] << 24);
            if ((uint64_t)p + 8 + sz > n) { rc = UVOL_ERR_TRUNCATED; break; }
            p += 8 + (size_t)sz; continue;
        }
        if (magic != 0xFD2FB528u) { rc = UVOL_ERR_CORRUPT; break; }
        p += 4;
        if (p >= n) { rc = UVOL_ERR_TRUNCATED; break; }
        const int fhd = src[p++];
        const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did = fhd & 3;
        if (fhd & 0x08) { rc = UVOL_ERR_CORRUPT; break; }            // reserved bit
        if (!single) { if (p >= n) { rc = UVOL_ERR_TRUNCATED; break; } p++; }       // window descriptor: the whole output is our window
        static const int did_bytes[4] = {0, 1, 2, 4};
        if (did) { uint32_t id = 0; if (p + did_bytes[did] > n) { rc = UVOL_ERR_TRUNCATED; break; } for (int i = 0; i < did_bytes[did]; i++) id |= (uint32_t)src[p + i] << (8 * i); p += did_bytes[did]; if (id) { rc = UVOL_ERR_UNSUPPORTED; break; } }
        const int fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
        if (p + fcs_bytes > n) { rc = UVOL_ERR_TRUNCATED; break; }
        uint64_t fcs = 0; for (int i = 0; i < fcs_bytes; i++) fcs |= (uint64_t)src[p + i] << (8 * i);
        if (fcs_bytes == 2) fcs += 256;
        p += fcs_bytes;
        const size_t frame_start = d;
        g_feat[15].fetch_add(1, std::memory_order_relaxed);
        FrameState fs;
        for (;;) {
            if (p + 3 > n) { rc = UVOL_ERR_TRUNCATED; break; }
            const uint32_t bh = src[p] | (src[p + 1] << 8) | ((uint32_t)src[p + 2] << 16); p += 3;
            const int last = bh & 1, type = (bh >> 1) & 3; const size_t bsize = bh >> 3;
            if (type < 3) g_feat[type].fetch_add(1, std::memory_order_relaxed);
            if (type == 0) { if (p + bsize > n) { rc = UVOL_ERR_TRUNCATED; break; } if (d + bsize > cap) { rc = UVOL_ERR_CORRUPT; break; } memcpy(dst + d, src + p, bsize); d += bsize; p += bsize; }
            else if (type == 1) { if (p + 1 > n) { rc = UVOL_ERR_TRUNCATED; break; } if (d + bsize > cap) { rc = UVOL_ERR_CORRUPT; break; } memset(dst + d, src[p], bsize); d += bsize; p += 1; }
            else if (type == 2) {
                if (p + bsize > n) { rc = UVOL_ERR_TRUNCATED; break; }
                size_t rel = d - frame_start;                        // offsets reach back to the start of the frame at most
                if (bsize > (128u << 10) || !block_decode(src + p, bsize, dst + frame_start, cap - frame_start, &rel, fs, lit_buf)) { rc = UVOL_ERR_CORRUPT; break; }
                d = frame_start + rel; p += bsize;
            } else { rc = UVOL_ERR_CORRUPT; break; }
            if (last) break;
        }
        if (rc) break;
        if (checksum) { if (p + 4 > n) { rc = UVOL_ERR_TRUNCATED; break; } p += 4; }
        if (fcs_bytes && d - frame_start != fcs) { rc = UVOL_ERR_CORRUPT; break; }
    }
    delete[] lit_buf;
    *out_len = d;
    return rc;
}
