// etc1s_encode.cpp -- minimal, deterministic KTX2 / BasisLZ (ETC1S, "video" texture array) ENCODER
// for synthetic inputs.  Bench/test INPUT tooling (no basisu binary exists in this image,
// SURVEY.md 7.2-3); not part of libuvol_b200.so, not part of oracle/.
//
// Output mirrors what scripts/Encoder.py:290 (`basisu -ktx2 -tex_type video -multifile_num B`)
// writes and the reference's fixtures contain: KTX2 container, vkFormat 0,
// supercompression 1 (BasisLZ), DFD colour model 163, KTXanimData key (video), one I-frame followed
// by P-frames whose unchanged blocks use the CR predictor; endpoint / selector codebooks, the four
// slice Huffman tables, selector history buffer, selector RLE and endpoint-predictor repeat runs.
// Quality is not a goal (the image content is synthetic); conformance of the bitstream is.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <queue>
#include <unordered_map>
#include <vector>

namespace {

typedef std::vector<uint8_t> Bytes;
struct BitW {
    Bytes b; uint64_t acc = 0; int n = 0;
    void put(uint32_t v, int bits) { if (!bits) return; acc |= (uint64_t)(v & (bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u))) << n; n += bits; while (n >= 8) { b.push_back((uint8_t)(acc & 255)); acc >>= 8; n -= 8; } }
    void flush() { if (n > 0) { b.push_back((uint8_t)(acc & 255)); acc = 0; n = 0; } }
    void vlc(uint32_t v, int cb) { for (;;) { uint32_t ch = v & ((1u << cb) - 1u); v >>= cb; if (v) ch |= 1u << cb; put(ch, cb + 1); if (!v) break; } }
};

// length-limited Huffman code lengths (heuristic: rebuild with flattened frequencies until it fits)
std::vector<uint8_t> huff_lengths(const std::vector<uint64_t> &freq_in, int maxlen) {
    const size_t n = freq_in.size(); std::vector<uint8_t> len(n, 0);
    std::vector<uint64_t> freq = freq_in;
    size_t used = 0; for (uint64_t f : freq) used += f != 0;
    if (used == 0) return len;
    if (used == 1) { for (size_t i = 0; i < n; i++) if (freq[i]) len[i] = 1; return len; }
    for (int iter = 0; iter < 40; iter++) {
        struct Node { uint64_t f; int l, r; };
        std::vector<Node> nodes; typedef std::pair<uint64_t, int> QE;
        std::priority_queue<QE, std::vector<QE>, std::greater<QE>> q;
        for (size_t i = 0; i < n; i++) if (freq[i]) { nodes.push_back({freq[i], -1, (int)i}); q.push({freq[i], (int)nodes.size() - 1}); }
        while (q.size() > 1) { QE a = q.top(); q.pop(); QE b2 = q.top(); q.pop(); nodes.push_back({a.first + b2.first, a.second, b2.second}); q.push({a.first + b2.first, (int)nodes.size() - 1}); }
        std::fill(len.begin(), len.end(), 0);
        int maxl = 0; std::vector<std::pair<int, int>> stk; stk.push_back({q.top().second, 0});
        while (!stk.empty()) { auto [id, d] = stk.back(); stk.pop_back(); const Node &nd = nodes[id]; if (nd.l < 0) { len[nd.r] = (uint8_t)(d ? d : 1); maxl = std::max(maxl, d); } else { stk.push_back({nd.l, d + 1}); stk.push_back({nd.r, d + 1}); } }
        if (maxl <= maxlen) return len;
        for (size_t i = 0; i < n; i++) if (freq[i]) freq[i] = (freq[i] + 1) / 2 + 1;     // flatten and retry
    }
    return len;
}
struct Code { std::vector<uint8_t> len; std::vector<uint32_t> rev; };   // rev = bit-reversed canonical code (LSB-first emission)
Code make_code(const std::vector<uint8_t> &len) {
    Code c; c.len = len; c.rev.assign(len.size(), 0);
    uint32_t cnt[18] = {0}, next[18] = {0}; for (uint8_t l : len) cnt[l]++;
    cnt[0] = 0; uint32_t code = 0; for (int l = 1; l <= 16; l++) { code = (code + cnt[l - 1]) << 1; next[l] = code; }
    for (size_t s = 0; s < len.size(); s++) if (len[s]) { uint32_t cd = next[len[s]]++, r = 0; for (int i = 0; i < len[s]; i++) r |= ((cd >> i) & 1u) << (len[s] - 1 - i); c.rev[s] = r; }
    return c;
}
void put_sym(BitW &w, const Code &c, uint32_t s) { w.put(c.rev[s], c.len[s]); }

// serialises a Huffman table (B.2): total 14b, ncl 5b, 3-bit code-length-code sizes, RLE'd sizes
const uint8_t CL_ORDER[21] = {17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16};
void write_table(BitW &w, const std::vector<uint8_t> &len) {
    size_t total = len.size(); while (total > 0 && len[total - 1] == 0) total--;
    w.put((uint32_t)total, 14);
    if (!total) return;
    std::vector<std::pair<uint8_t, uint32_t>> toks;       // (cl symbol, extra value)
    for (size_t i = 0; i < total;) {
        if (len[i] == 0) { size_t run = 0; while (i + run < total && len[i + run] == 0) run++;
            size_t left = run; while (left >= 11) { size_t r = std::min<size_t>(left, 138); toks.push_back({18, (uint32_t)(r - 11)}); left -= r; }
            if (left >= 3) { toks.push_back({17, (uint32_t)(left - 3)}); left = 0; }
            while (left--) toks.push_back({0, 0});
            i += run; }
        else { size_t run = 1; while (i + run < total && len[i + run] == len[i]) run++;
            toks.push_back({len[i], 0}); size_t left = run - 1;
            while (left >= 7) { size_t r = std::min<size_t>(left, 134); toks.push_back({20, (uint32_t)(r - 7)}); left -= r; }
            if (left >= 3) { toks.push_back({19, (uint32_t)(left - 3)}); left = 0; }
            while (left--) toks.push_back({len[i], 0});
            i += run; }
    }
    std::vector<uint64_t> f(21, 0); for (auto &t : toks) f[t.first]++;
    const std::vector<uint8_t> cl = huff_lengths(f, 7); const Code cc = make_code(cl);
    int ncl = 21; while (ncl > 1 && cl[CL_ORDER[ncl - 1]] == 0) ncl--;
    w.put((uint32_t)ncl, 5);
    for (int i = 0; i < ncl; i++) w.put(cl[CL_ORDER[i]], 3);
    for (auto &t : toks) { put_sym(w, cc, t.first); if (t.first == 17) w.put(t.second, 3); else if (t.first == 18) w.put(t.second, 7); else if (t.first == 19) w.put(t.second, 2); else if (t.first == 20) w.put(t.second, 7); }
}

const int INTEN[8][4] = {{-8, -2, 2, 8}, {-17, -5, 5, 17}, {-29, -9, 9, 29}, {-42, -13, 13, 42}, {-60, -18, 18, 60}, {-80, -24, 24, 80}, {-106, -33, 33, 106}, {-183, -47, 47, 183}};

void put32(Bytes &b, uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((v >> (8 * i)) & 255); }
void put64(Bytes &b, uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((v >> (8 * i)) & 255); }
void put16(Bytes &b, uint32_t v) { b.push_back(v & 255); b.push_back((v >> 8) & 255); }
void pad_to(Bytes &b, size_t a) { while (b.size() % a) b.push_back(0); }

}  // namespace

// rgba: layers * h * w * 4 bytes (alpha ignored).  Returns malloc'd .ktx2 bytes.
extern "C" size_t uvsynth_etc1s_encode(const uint8_t *rgba, uint32_t w, uint32_t h, uint32_t layers, int max_endpoints, uint8_t **out_buf) {
    *out_buf = nullptr;
    if (!w || !h || !layers || (w & 3) || (h & 3)) return 0;
    const uint32_t bx = w / 4, by = h / 4, nblk = bx * by;
    // ---- per block: base colour (5:5:5), intensity table, plane-fit selector pattern
    std::vector<uint32_t> ep_raw((size_t)layers * nblk), sel_raw((size_t)layers * nblk);
    for (uint32_t L = 0; L < layers; L++) for (uint32_t yb = 0; yb < by; yb++) for (uint32_t xb = 0; xb < bx; xb++) {
        int sum[3] = {0, 0, 0}; int lum[16];
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) { const uint8_t *p = rgba + (((size_t)L * h + yb * 4 + y) * w + xb * 4 + x) * 4; sum[0] += p[0]; sum[1] += p[1]; sum[2] += p[2]; lum[y * 4 + x] = p[0] + 2 * p[1] + p[2]; }
        int c5[3]; for (int c = 0; c < 3; c++) { c5[c] = (sum[c] / 16 * 31 + 127) / 255; c5[c] = std::min(31, std::max(0, c5[c])); }
        int avg = 0; for (int i = 0; i < 16; i++) avg += lum[i]; avg /= 16;
        int maxd = 0; for (int i = 0; i < 16; i++) maxd = std::max(maxd, abs(lum[i] - avg) / 4);
        int inten = 0; for (int t = 0; t < 8; t++) if (abs(INTEN[t][3] - maxd) < abs(INTEN[inten][3] - maxd)) inten = t;
        // least-squares plane d(x,y) = a*(x-1.5) + b*(y-1.5) in luma, quantised -> selector ramp
        double sa = 0, sb = 0; for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) { const double d = (lum[y * 4 + x] - avg) / 4.0; sa += d * (x - 1.5); sb += d * (y - 1.5); }
        const double scale = INTEN[inten][3] > 0 ? 1.0 / INTEN[inten][3] : 0.0;
        int qa = (int)lrint(sa / 20.0 * scale * 6.0), qb = (int)lrint(sb / 20.0 * scale * 6.0);    // slope in 1/6 "outer modifier" per pixel
        qa = std::min(6, std::max(-6, qa)); qb = std::min(6, std::max(-6, qb));
        // offset term: where the block mean sits relative to the quantised base colour
        const int base_l = ((c5[0] << 3 | c5[0] >> 2) + 2 * (c5[1] << 3 | c5[1] >> 2) + (c5[2] << 3 | c5[2] >> 2));
        int qc = (int)lrint((avg - base_l) / 4.0 * scale * 2.0); qc = std::min(1, std::max(-1, qc));
        uint32_t sel = 0;
        for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
            const double v = (qa * (x - 1.5) + qb * (y - 1.5)) / 6.0 + qc * 0.5;       // in units of the outer modifier
            int s = v < -0.55 ? 0 : (v < 0 ? 1 : (v < 0.55 ? 2 : 3));
            sel |= (uint32_t)s << (8 * y + 2 * x);
        }
        ep_raw[(size_t)L * nblk + yb * bx + xb] = (uint32_t)c5[0] | (c5[1] << 8) | (c5[2] << 16) | ((uint32_t)inten << 24);
        sel_raw[(size_t)L * nblk + yb * bx + xb] = sel;
    }
    // ---- codebooks.  Endpoints: exact colours if they fit, else drop low colour bits until they do.
    std::vector<uint32_t> epal, spal; std::vector<uint16_t> eidx(ep_raw.size()), sidx(sel_raw.size());
    for (int shift = 0; shift < 5; shift++) {
        std::unordered_map<uint32_t, uint32_t> m; epal.clear(); bool ok = true;
        for (size_t i = 0; i < ep_raw.size(); i++) {
            uint32_t e = ep_raw[i], q = 0;
            for (int c = 0; c < 3; c++) { uint32_t v = (e >> (8 * c)) & 31; v = (v >> shift) << shift; if (shift) v |= v >> (5 - shift); q |= std::min(31u, v) << (8 * c); }
            q |= e & 0xff000000u;
            auto it = m.find(q);
            if (it == m.end()) { if ((int)epal.size() >= max_endpoints) { ok = false; break; } it = m.emplace(q, (uint32_t)epal.size()).first; epal.push_back(q); }
            eidx[i] = (uint16_t)it->second;
        }
        if (ok) break;
        if (shift == 4) return 0;
    }
    { std::unordered_map<uint32_t, uint32_t> m; for (size_t i = 0; i < sel_raw.size(); i++) { auto it = m.find(sel_raw[i]); if (it == m.end()) { it = m.emplace(sel_raw[i], (uint32_t)spal.size()).first; spal.push_back(sel_raw[i]); } sidx[i] = (uint16_t)it->second; } }
    const uint32_t ec = (uint32_t)epal.size(), scnt = (uint32_t)spal.size(), HIST = 64;
    if (ec > 16128 || scnt > 16128) return 0;
    // ---- endpoint codebook blob
    BitW we;
    {
        std::vector<uint64_t> f0(32, 0), f1(32, 0), f2(32, 0), fi(8, 0);
        uint32_t prev[3] = {16, 16, 16}, pint = 0;
        for (uint32_t i = 0; i < ec; i++) { const uint32_t e = epal[i]; fi[((e >> 24) - pint) & 7]++; pint = e >> 24;
            for (int c = 0; c < 3; c++) { const uint32_t v = (e >> (8 * c)) & 31, d = (v - prev[c]) & 31; (prev[c] <= 9 ? f0 : prev[c] <= 21 ? f1 : f2)[d]++; prev[c] = v; } }
        const Code c0 = make_code(huff_lengths(f0, 16)), c1 = make_code(huff_lengths(f1, 16)), c2 = make_code(huff_lengths(f2, 16)), ci = make_code(huff_lengths(fi, 16));
        write_table(we, c0.len); write_table(we, c1.len); write_table(we, c2.len); write_table(we, ci.len);
        we.put(0, 1);   // not grayscale
        prev[0] = prev[1] = prev[2] = 16; pint = 0;
        for (uint32_t i = 0; i < ec; i++) { const uint32_t e = epal[i]; put_sym(we, ci, ((e >> 24) - pint) & 7); pint = e >> 24;
            for (int c = 0; c < 3; c++) { const uint32_t v = (e >> (8 * c)) & 31, d = (v - prev[c]) & 31; put_sym(we, prev[c] <= 9 ? c0 : prev[c] <= 21 ? c1 : c2, d); prev[c] = v; } }
        we.flush();
    }
    // ---- selector codebook blob
    BitW ws;
    {
        ws.put(0, 1); ws.put(0, 1); ws.put(0, 1);
        std::vector<uint64_t> f(256, 0); uint32_t prev = 0;
        for (uint32_t i = 1; i < scnt; i++) { prev = spal[i - 1]; for (int j = 0; j < 4; j++) f[((spal[i] ^ prev) >> (8 * j)) & 255]++; }
        const Code c = make_code(huff_lengths(f, 16));
        write_table(ws, c.len);
        for (uint32_t i = 0; i < scnt; i++) for (int j = 0; j < 4; j++) { if (i == 0) ws.put((spal[0] >> (8 * j)) & 255, 8); else put_sym(ws, c, ((spal[i] ^ spal[i - 1]) >> (8 * j)) & 255); }
        ws.flush();
    }
    // ---- slices: first pass builds the symbol streams, second pass emits bits
    struct Tok { uint8_t kind; uint32_t v; };    // 0 endpoint_pred sym, 1 vlc4 after pred repeat, 2 delta_endpoint, 3 selector sym, 4 rle sym, 5 vlc7
    std::vector<std::vector<Tok>> toks(layers);
    std::vector<uint64_t> f_epm(257, 0), f_dem(ec, 0), f_sm(scnt + HIST + 1, 0), f_rle(64, 0);
    for (uint32_t L = 0; L < layers; L++) {
        const uint16_t *E = eidx.data() + (size_t)L * nblk, *S = sidx.data() + (size_t)L * nblk;
        const uint16_t *PE = L ? E - nblk : nullptr, *PS = L ? S - nblk : nullptr;
        std::vector<uint8_t> pred(nblk); uint32_t prev_ep = 0;
        for (uint32_t y = 0; y < by; y++) for (uint32_t x = 0; x < bx; x++) {
            const uint32_t bi = y * bx + x; uint8_t p;
            if (PE && PE[bi] == E[bi] && PS[bi] == S[bi]) p = 2;
            else if (x > 0 && E[bi] == prev_ep) p = 0;
            else if (y > 0 && E[bi] == E[bi - bx]) p = 1;
            else p = 3;
            pred[bi] = p; prev_ep = E[bi];
        }
        std::vector<Tok> &T = toks[L];
        uint32_t hist[HIST]; memset(hist, 0, sizeof hist); uint32_t rover = HIST / 2; prev_ep = 0;
        uint32_t prev_sym = 0; int pending_rep = 0; bool have_prev = false;
        // pre-compute the 2x2 predictor symbols in decode order to find repeat runs
        std::vector<uint32_t> psyms; for (uint32_t y = 0; y < by; y += 2) for (uint32_t x = 0; x < bx; x += 2) {
            auto P = [&](uint32_t xx, uint32_t yy) -> uint32_t { return (xx < bx && yy < by) ? pred[yy * bx + xx] : 0u; };
            psyms.push_back(P(x, y) | (P(x + 1, y) << 2) | (P(x, y + 1) << 4) | (P(x + 1, y + 1) << 6)); }
        size_t psi = 0; int rle_left = 0;
        // selectors of non-CR blocks in decode order, to find RLE runs
        for (uint32_t y = 0; y < by; y++) for (uint32_t x = 0; x < bx; x++) {
            const uint32_t bi = y * bx + x;
            if ((x & 1) == 0 && (y & 1) == 0) {
                const uint32_t s = psyms[psi];
                if (pending_rep > 0) pending_rep--;
                else {
                    size_t run = 0; if (have_prev && s == prev_sym) { while (psi + run < psyms.size() && psyms[psi + run] == prev_sym && run < 3 + 255) run++; }
                    if (run >= 3) { T.push_back({0, 256}); f_epm[256]++; T.push_back({1, (uint32_t)(run - 3)}); pending_rep = (int)run - 1; }
                    else { T.push_back({0, s}); f_epm[s]++; prev_sym = s; have_prev = true; }
                }
                psi++;
            }
            const uint8_t p = pred[bi];
            if (p == 3) { const uint32_t d = (E[bi] + ec - prev_ep) % ec; T.push_back({2, d}); f_dem[d]++; }
            prev_ep = E[bi];
            if (p == 2) continue;
            const uint32_t s = S[bi];
            if (rle_left > 0) { rle_left--; continue; }
            // RLE of hist[0]
            if (s == hist[0]) {
                size_t run = 0; { uint32_t yy = y, xx = x; while (run < 3 + 62) { const uint32_t b2 = yy * bx + xx; if (pred[b2] != 2) { if (S[b2] != hist[0]) break; run++; } if (++xx == bx) { xx = 0; if (++yy == by) break; } } }
                if (run >= 3) { T.push_back({3, scnt + HIST}); f_sm[scnt + HIST]++; T.push_back({4, (uint32_t)(run - 3)}); f_rle[run - 3]++; rle_left = (int)run - 1; continue; }
            }
            int hi = -1; for (uint32_t i = 0; i < HIST; i++) if (hist[i] == s) { hi = (int)i; break; }
            if (hi >= 0) { T.push_back({3, scnt + (uint32_t)hi}); f_sm[scnt + hi]++; if (hi) std::swap(hist[hi / 2], hist[hi]); }
            else { T.push_back({3, s}); f_sm[s]++; hist[rover++] = s; if (rover == HIST) rover = HIST / 2; }
        }
    }
    const Code c_epm = make_code(huff_lengths(f_epm, 16)), c_dem = make_code(huff_lengths(f_dem, 16)), c_sm = make_code(huff_lengths(f_sm, 16)), c_rle = make_code(huff_lengths(f_rle, 16));
    BitW wt; write_table(wt, c_epm.len); write_table(wt, c_dem.len); write_table(wt, c_sm.len); write_table(wt, c_rle.len); wt.put(HIST, 13); wt.flush();
    std::vector<Bytes> slice(layers);
    for (uint32_t L = 0; L < layers; L++) {
        BitW w2;
        for (const Tok &t : toks[L]) {
            if (t.kind == 0) put_sym(w2, c_epm, t.v); else if (t.kind == 1) w2.vlc(t.v, 4); else if (t.kind == 2) put_sym(w2, c_dem, t.v);
            else if (t.kind == 3) put_sym(w2, c_sm, t.v); else if (t.kind == 4) put_sym(w2, c_rle, t.v); else w2.vlc(t.v, 7);
        }
        w2.flush(); slice[L] = w2.b;
    }
    // ---- KTX2 container (B.1)
    Bytes k; const uint8_t id[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x32, 0x30, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};
    k.insert(k.end(), id, id + 12);
    put32(k, 0); put32(k, 1); put32(k, w); put32(k, h); put32(k, 0); put32(k, layers); put32(k, 1); put32(k, 1); put32(k, 1);
    Bytes kvd;
    { const char key1[] = "KTXanimData"; Bytes v; put32(v, 1); put32(v, 15); put32(v, 0);
      put32(kvd, (uint32_t)(sizeof key1 + v.size())); kvd.insert(kvd.end(), key1, key1 + sizeof key1); kvd.insert(kvd.end(), v.begin(), v.end()); pad_to(kvd, 4);
      const char key2[] = "KTXwriter"; const char val2[] = "uvol-b200 synth (BasisLZ/ETC1S video)";
      put32(kvd, (uint32_t)(sizeof key2 + sizeof val2)); kvd.insert(kvd.end(), key2, key2 + sizeof key2); kvd.insert(kvd.end(), val2, val2 + sizeof val2); pad_to(kvd, 4); }
    Bytes sgd; put16(sgd, ec); put16(sgd, scnt); put32(sgd, (uint32_t)we.b.size()); put32(sgd, (uint32_t)ws.b.size()); put32(sgd, (uint32_t)wt.b.size()); put32(sgd, 0);
    uint32_t off = 0; for (uint32_t L = 0; L < layers; L++) { put32(sgd, L ? 2u : 0u); put32(sgd, off); put32(sgd, (uint32_t)slice[L].size()); put32(sgd, 0); put32(sgd, 0); off += (uint32_t)slice[L].size(); }
    sgd.insert(sgd.end(), we.b.begin(), we.b.end()); sgd.insert(sgd.end(), ws.b.begin(), ws.b.end()); sgd.insert(sgd.end(), wt.b.begin(), wt.b.end());
    const uint32_t dfdOff = 80 + 24, dfdLen = 44, kvdOff = dfdOff + dfdLen, kvdLen = (uint32_t)kvd.size();
    const uint64_t sgdOff = (kvdOff + kvdLen + 7) / 8 * 8, sgdLen = sgd.size(), lvOff = sgdOff + sgdLen, lvLen = off;
    put32(k, dfdOff); put32(k, dfdLen); put32(k, kvdOff); put32(k, kvdLen); put64(k, sgdOff); put64(k, sgdLen);
    put64(k, lvOff); put64(k, lvLen); put64(k, 0);
    put32(k, 44); put32(k, 0); put16(k, 2); put16(k, 40); k.push_back(163); k.push_back(1); k.push_back(2); k.push_back(0);
    k.push_back(3); k.push_back(3); k.push_back(0); k.push_back(0); for (int i = 0; i < 8; i++) k.push_back(0);
    put16(k, 0); k.push_back(63); k.push_back(0); for (int i = 0; i < 4; i++) k.push_back(0); put32(k, 0); put32(k, 0xFFFFFFFFu);
    k.insert(k.end(), kvd.begin(), kvd.end());
    while (k.size() < sgdOff) k.push_back(0);
    k.insert(k.end(), sgd.begin(), sgd.end());
    for (uint32_t L = 0; L < layers; L++) k.insert(k.end(), slice[L].begin(), slice[L].end());
    *out_buf = (uint8_t *)malloc(k.size() + 16); memcpy(*out_buf, k.data(), k.size());
    return k.size();
}
