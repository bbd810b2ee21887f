// draco_encode.cpp -- minimal, deterministic Draco 2.2 mesh ENCODER for synthetic inputs.
//
// Why it exists: the benchmark configs of BASELINE.json need 50k / 200k-vertex frames, the reference
// ships only 26k-vertex fixtures, and no draco_encoder binary exists in this image.
// It is bench/test INPUT tooling: it is not part of libuvol_b200.so and not part of oracle/.
//
// What it writes (same feature set as the reference's fixtures, SURVEY.md A.1, which is what
// scripts/Encoder.py:260 "draco_encoder -qp 11 -qt 10 -qn 8 -cl 7" produces): bitstream 2.2,
// edgebreaker with VALENCE traversal, POSITION (vertex attribute, PARALLELOGRAM+WRAP, quantised),
// TEX_COORD (corner attribute, TEX_COORDS_PORTABLE+WRAP, quantised), NORMAL (corner attribute,
// GEOMETRIC_NORMAL + canonicalised octahedron), every symbol run RAW rANS, bit runs rABS.
// Input restriction: one closed, manifold, genus-0 component (no boundary, no handles).
//
// Method: "decoder in the loop".  The connectivity symbols come from an edgebreaker traversal of
// the input; everything order-dependent (valence contexts, seam-bit order, attribute vertex ids,
// entry order, predictions) is obtained by running the product's own host-callable decode logic
// (csrc/draco_core.h) on the connectivity just produced, so the encoder emits exactly the
// corrections the decoder will undo.  Tests close the loop independently: oracle(decode(encode(m)))
// must reproduce m (tests/test_synth.py).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <unordered_map>
#include <vector>
#include "../../universal-volumetric_b200/csrc/uvol_internal.h"
#include "../../universal-volumetric_b200/csrc/draco_core.h"

namespace {

typedef std::vector<uint8_t> Bytes;
void put_u8(Bytes &b, uint32_t v) { b.push_back((uint8_t)v); }
void put_u16(Bytes &b, uint32_t v) { put_u8(b, v & 255); put_u8(b, v >> 8); }
void put_u32(Bytes &b, uint32_t v) { for (int i = 0; i < 4; i++) put_u8(b, (v >> (8 * i)) & 255); }
void put_f32(Bytes &b, float f) { uint32_t v; memcpy(&v, &f, 4); put_u32(b, v); }
void put_varint(Bytes &b, uint64_t v) { while (v >= 128) { put_u8(b, (uint32_t)(v & 127) | 128); v >>= 7; } put_u8(b, (uint32_t)v); }
void put_bytes(Bytes &b, const Bytes &s) { b.insert(b.end(), s.begin(), s.end()); }

// ---- rABS bit run: prob_zero u8, size varint, data (A.2)
void write_rabs(Bytes &out, const std::vector<uint8_t> &bits) {
    uint64_t zeros = 0; for (uint8_t x : bits) zeros += !x;
    const uint64_t total = bits.size() ? bits.size() : 1;
    uint32_t raw = (uint32_t)(((double)zeros / (double)total) * 256.0 + 0.5);
    uint32_t p0 = raw < 255 ? raw : 255; if (p0 == 0) p0 = 1;
    const uint32_t p = 256 - p0;
    Bytes buf; uint32_t state = 4096;
    for (size_t i = bits.size(); i-- > 0;) {
        const int val = bits[i]; const uint32_t ls = val ? p : p0;
        if (state >= 4096u / 256u * 256u * ls) { buf.push_back((uint8_t)(state & 255)); state >>= 8; }
        const uint32_t q = state / ls, rem = state - q * ls;
        state = q * 256 + rem + (val ? 0 : p);
    }
    state -= 4096;
    if (state < (1u << 6)) buf.push_back((uint8_t)state);
    else if (state < (1u << 14)) { const uint32_t v = (1u << 14) + state; buf.push_back(v & 255); buf.push_back(v >> 8); }
    else { const uint32_t v = (2u << 22) + state; buf.push_back(v & 255); buf.push_back((v >> 8) & 255); buf.push_back(v >> 16); }
    put_u8(out, p0); put_varint(out, buf.size()); put_bytes(out, buf);
}

// ---- rANS symbol run body: probability table (alphabet = max symbol + 1, normalised to 2^pb), size varint, data (A.2)
void write_rans_run(Bytes &out, const std::vector<uint32_t> &syms, int pb) {
    uint32_t maxs = 0; for (uint32_t s : syms) maxs = std::max(maxs, s);
    const uint32_t prec = 1u << pb, A = maxs + 1;
    std::vector<uint64_t> freq(A, 0); for (uint32_t s : syms) freq[s]++;
    // normalise to `prec` keeping every used symbol >= 1
    std::vector<uint32_t> prob(A, 0); uint64_t tot = syms.size() ? syms.size() : 1; int64_t sum = 0;
    for (uint32_t i = 0; i < A; i++) if (freq[i]) { uint64_t p = freq[i] * prec / tot; if (p == 0) p = 1; prob[i] = (uint32_t)p; sum += (int64_t)p; }
    if (syms.empty()) { prob[0] = prec; sum = prec; }
    // fix the rounding error on the most probable symbols
    std::vector<uint32_t> order(A); for (uint32_t i = 0; i < A; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return prob[a] != prob[b] ? prob[a] > prob[b] : a < b; });
    int64_t err = (int64_t)prec - sum;
    while (err != 0) {
        bool moved = false;
        for (uint32_t k = 0; k < A && err != 0; k++) {
            const uint32_t i = order[k]; if (!prob[i]) continue;
            if (err > 0) { prob[i]++; err--; moved = true; }
            else if (prob[i] > 1) { prob[i]--; err++; moved = true; }
        }
        if (!moved) break;
    }
    std::vector<uint32_t> cum(A + 1, 0); for (uint32_t i = 0; i < A; i++) cum[i + 1] = cum[i] + prob[i];
    put_varint(out, A);
    for (uint32_t i = 0; i < A;) {
        if (prob[i] == 0) { uint32_t run = 0; while (i + run < A && prob[i + run] == 0 && run < 64) run++; put_u8(out, ((run - 1) << 2) | 3); i += run; }
        else {
            const uint32_t p = prob[i]; const int extra = p < (1u << 6) ? 0 : (p < (1u << 14) ? 1 : 2);
            put_u8(out, ((p & 63) << 2) | (uint32_t)extra);
            for (int k = 0; k < extra; k++) put_u8(out, (p >> (8 * (k + 1) - 2)) & 255);
            i++;
        }
    }
    const uint32_t lbase = prec * 4;
    Bytes buf; uint64_t state = lbase;
    for (size_t i = syms.size(); i-- > 0;) {
        const uint32_t s = syms[i], p = prob[s];
        while (state >= (uint64_t)(lbase / prec) * 256ull * p) { buf.push_back((uint8_t)(state & 255)); state >>= 8; }
        state = (state / p) * prec + state % p + cum[s];
    }
    state -= lbase;
    if (state < (1u << 6)) buf.push_back((uint8_t)state);
    else if (state < (1u << 14)) { const uint32_t v = (1u << 14) + (uint32_t)state; buf.push_back(v & 255); buf.push_back(v >> 8); }
    else if (state < (1u << 22)) { const uint32_t v = (2u << 22) + (uint32_t)state; buf.push_back(v & 255); buf.push_back((v >> 8) & 255); buf.push_back(v >> 16); }
    else { const uint32_t v = (3u << 30) + (uint32_t)state; for (int k = 0; k < 4; k++) buf.push_back((v >> (8 * k)) & 255); }
    put_varint(out, buf.size()); put_bytes(out, buf);
}

// ---- RAW scheme: scheme u8=1, max_bit_length u8, then the run with precision (3 * max_bit_length) / 2 clamped to [12, 20]
void write_symbols_raw(Bytes &out, const std::vector<uint32_t> &syms) {
    uint32_t maxs = 0; for (uint32_t s : syms) maxs = std::max(maxs, s);
    int mbl = 1; while ((1u << mbl) <= maxs) mbl++;
    int pb = (3 * mbl) / 2; pb = pb < 12 ? 12 : (pb > 20 ? 20 : pb);
    put_u8(out, 1); put_u8(out, (uint32_t)mbl);
    write_rans_run(out, syms, pb);
}

// ---- TAGGED scheme (draco::EncodeTaggedSymbols): scheme u8=0, a 12-bit-precision run of one TAG per value tuple (the bit length
// that holds every component of the tuple, at least 1), then the components themselves as raw LSB-first bit fields, byte padded.
void write_symbols_tagged(Bytes &out, const std::vector<uint32_t> &syms, int nc) {
    std::vector<uint32_t> tags(syms.size() / (size_t)nc);
    for (size_t i = 0; i < tags.size(); i++) {
        uint32_t m = 1; for (int c = 0; c < nc; c++) m = std::max(m, syms[i * nc + c]);
        int bl = 1; while (bl < 32 && (m >> bl) != 0) bl++;
        tags[i] = (uint32_t)bl;
    }
    put_u8(out, 0);
    write_rans_run(out, tags, 12);
    Bytes bits; uint64_t acc = 0; int have = 0;
    for (size_t i = 0; i < tags.size(); i++) for (int c = 0; c < nc; c++) {
        acc |= (uint64_t)syms[i * nc + c] << have; have += (int)tags[i];
        while (have >= 8) { bits.push_back((uint8_t)(acc & 255)); acc >>= 8; have -= 8; }
    }
    if (have > 0) bits.push_back((uint8_t)(acc & 255));
    put_bytes(out, bits);
}
// which attributes use the TAGGED scheme (bit 0 position, 1 uv, 2 normal): a generator switch for the parity tests
int g_tagged_mask = 0;
void write_symbols(Bytes &out, const std::vector<uint32_t> &syms, int nc, int which) {
    if (g_tagged_mask & (1 << which)) write_symbols_tagged(out, syms, nc); else write_symbols_raw(out, syms);
}

uint32_t zigzag_sym(int32_t v) { return v >= 0 ? (uint32_t)v << 1 : (((uint32_t)(-(v + 1))) << 1) | 1u; }

int32_t wrap_corr(int32_t orig, long long pred, int32_t mn, int32_t mx) {
    if (pred > mx) pred = mx;
    if (pred < mn) pred = mn;
    const int32_t md = 1 + mx - mn; int32_t c = orig - (int32_t)pred;
    if (c > md / 2) c -= md; else if (c < -(md / 2)) c += md;
    return c;
}

}  // namespace

// Encodes one mesh.  pos[nv*3]; faces_pos[nf*3] position indices; uv[nuv*2], faces_uv[nf*3] per-corner uv
// indices; nrm[nv*3] per-position-vertex normals.  Returns a malloc'd buffer (caller frees with
// uvsynth_free) and its size, or 0 on failure (non-manifold / open / multi-component input).
extern "C" size_t uvsynth_draco_encode(const float *pos, uint32_t nv, const uint32_t *faces_pos, uint32_t nf, const float *uv, uint32_t nuv,
                                       const uint32_t *faces_uv, const float *nrm, int qp, int qt, int qn, uint8_t **out_buf) {
    *out_buf = nullptr;
    const int F = (int)nf, C = 3 * F;
    // ---- input corner table
    std::vector<int> oopp(C, -1);
    {
        std::unordered_map<uint64_t, int> edge; edge.reserve((size_t)C * 2);
        for (int c = 0; c < C; c++) {   // corner c is opposite the directed edge next(c) -> prev(c)
            const uint64_t a = faces_pos[cnext(c)], b = faces_pos[cprev(c)];
            if (!edge.emplace((a << 32) | b, c).second) return 0;
        }
        for (int c = 0; c < C; c++) {
            const uint64_t a = faces_pos[cnext(c)], b = faces_pos[cprev(c)];
            auto it = edge.find((b << 32) | a);
            if (it == edge.end()) return 0;                  // boundary: not supported
            oopp[c] = it->second;
        }
    }
    // ---- edgebreaker traversal (closed, genus 0): symbols in ENCODER order
    std::vector<uint8_t> sym; std::vector<int> sym_corner; sym.reserve(F); sym_corner.reserve(F);
    std::vector<uint8_t> fvis(F, 0), vvis(nv, 0);
    const int start = 0;   // init face = face 0, interior configuration
    fvis[0] = 1; vvis[faces_pos[0]] = vvis[faces_pos[1]] = vvis[faces_pos[2]] = 1;
    {
        std::vector<int> st; st.push_back(oopp[cnext(start)]);
        while (!st.empty()) {
            int c = st.back();
            if (c < 0 || fvis[c / 3]) { st.pop_back(); continue; }
            for (;;) {
                fvis[c / 3] = 1; sym_corner.push_back(c);
                const uint32_t v = faces_pos[c];
                if (!vvis[v]) { vvis[v] = 1; sym.push_back(0); c = oopp[cnext(c)]; continue; }           // C
                const int rc = oopp[cnext(c)], lc = oopp[cprev(c)];
                const bool rv = fvis[rc / 3], lv = fvis[lc / 3];
                if (rv) {
                    if (lv) { sym.push_back(4); st.pop_back(); break; }                                     // E
                    sym.push_back(3); c = lc;                                                               // R
                } else {
                    if (lv) { sym.push_back(2); c = rc; }                                                   // L
                    else { sym.push_back(1); st.back() = lc; st.push_back(rc); break; }                     // S
                }
            }
        }
    }
    const int nsym = (int)sym.size();
    if (nsym != F - 1) return 0;                             // more than one component (or a handle): not supported
    for (uint32_t v = 0; v < nv; v++) if (!vvis[v]) return 0;
    int nsplit = 0; for (uint8_t s : sym) nsplit += s == 1;
    // ---- run the decoder's connectivity on these symbols (decode order = reversed)
    std::vector<uint8_t> dsym(nsym); for (int i = 0; i < nsym; i++) dsym[i] = sym[nsym - 1 - i];
    DracoFrame f; memset(&f, 0, sizeof f);
    f.trav = 2; f.nv_enc = nv; f.nf = nf; f.nad = 2; f.nsym = (uint32_t)nsym; f.nsplit = (uint32_t)nsplit; f.nts = 0;
    Bytes sf_run; { std::vector<uint8_t> bits(1, 1); write_rabs(sf_run, bits); }
    {   // locate the run inside its own buffer for eb_decode_frame (prob u8, varint size (1 byte here), data)
        f.start_faces.prob_zero = sf_run[0]; f.start_faces.data_len = sf_run[1]; f.start_faces.data_off = 2;
    }
    const int maxv = (int)nv + nsplit + 4;
    std::vector<int> opp(C), c2v(C), lmc(maxv), val(maxv), stack(nsym + 16), invalid(nsplit + 8), skey(4), sval(4);
    std::vector<uint8_t> hole(maxv); std::vector<int8_t> ctx_log(nsym + 1);
    EbMem m; m.opp = opp.data(); m.c2v = c2v.data(); m.lmc = lmc.data(); m.val = val.data(); m.stack = stack.data(); m.skey = skey.data(); m.sval = sval.data();
    m.invalid = invalid.data(); m.hole = hole.data(); for (int i = 0; i < 6; i++) m.ctxsym[i] = nullptr;
    m.force_syms = dsym.data(); m.ctx_log = ctx_log.data();
    uint32_t aux0 = 0, slots = 0;
    if (eb_decode_frame(f, sf_run.data(), &aux0, m, &slots)) return 0;
    const int V = (int)slots;
    // decoder corner -> input corner (derivation in the header comment of tools/synth/README)
    std::vector<int> d2o(C);
    for (int sid = 0; sid < nsym; sid++) { const int c = sym_corner[nsym - 1 - sid]; d2o[3 * sid] = c; d2o[3 * sid + 1] = cnext(c); d2o[3 * sid + 2] = cprev(c); }
    d2o[3 * nsym] = cnext(start); d2o[3 * nsym + 1] = cprev(start); d2o[3 * nsym + 2] = start;
    {   // consistency: decoder vertices <-> input vertices must be a bijection, adjacency must agree
        std::vector<int> vmap(V, -1);
        for (int c = 0; c < C; c++) { const int dv = c2v[c], ov = (int)faces_pos[d2o[c]]; if (vmap[dv] < 0) vmap[dv] = ov; else if (vmap[dv] != ov) return 0; }
        for (int c = 0; c < C; c++) if (opp[c] < 0 || d2o[opp[c]] != oopp[d2o[c]]) return 0;
    }
    // ---- valence context arrays: context c is read from the back in decode order
    std::vector<std::vector<uint32_t>> ctxs(6);
    for (int sid = nsym - 1; sid >= 1; sid--) ctxs[ctx_log[sid]].push_back(dsym[sid]);   // sid 0 is the implicit E
    if (nsym > 0 && dsym[0] != 4) return 0;
    // ---- seams (decoder order), attribute corner tables, points, traversals: product logic on the host
    std::vector<uint8_t> eos[2], vos[2]; std::vector<int> ac2v[2], afirst[2], acnt[2];
    std::vector<uint8_t> seambits[2];
    for (int i = 0; i < 2; i++) { eos[i].assign(C, 0); vos[i].assign(V + 1, 0); ac2v[i].assign(C, -1); afirst[i].assign(V + 1, 0); acnt[i].assign(V + 1, 0); }
    for (int c = 0; c < C; c++) {
        const int o = opp[c];
        if (o / 3 < c / 3) continue;
        const int oc = d2o[c], oo = d2o[o];
        const bool uvseam = faces_uv[cnext(oc)] != faces_uv[cprev(oo)] || faces_uv[cprev(oc)] != faces_uv[cnext(oo)];
        seambits[0].push_back(uvseam); seambits[1].push_back(0);
        if (uvseam) seam_mark(c, opp.data(), c2v.data(), eos[0].data(), vos[0].data());
    }
    int err = 0; uint32_t attr_verts[2];
    for (int i = 0; i < 2; i++) {
        for (int v = 0; v < V; v++) acnt[i][v] = attr_vertex_fan(v, opp.data(), lmc.data(), eos[i].data(), vos[i].data(), afirst[i].data(), ac2v[i].data(), 0, 0, F, &err);
        int run = 0; for (int v = 0; v < V; v++) { const int c = acnt[i][v]; acnt[i][v] = run; run += c; }
        attr_verts[i] = (uint32_t)run;
        for (int v = 0; v < V; v++) attr_vertex_fan(v, opp.data(), lmc.data(), eos[i].data(), vos[i].data(), afirst[i].data(), ac2v[i].data(), acnt[i][v], 1, F, &err);
    }
    if (err) return 0;
    TableView tv[3];
    tv[0] = TableView{opp.data(), c2v.data(), nullptr, nullptr, nullptr};
    for (int i = 0; i < 2; i++) tv[1 + i] = TableView{opp.data(), c2v.data(), eos[i].data(), ac2v[i].data(), vos[i].data()};
    std::vector<int> d2c[3], v2d1[3]; uint32_t entries[3];
    for (int t = 0; t < 3; t++) {
        const int nvt = t == 0 ? V : (int)attr_verts[t - 1];
        d2c[t].assign(nvt + 4, 0); v2d1[t].assign(nvt + 4, 0);
        std::vector<uint8_t> fv(F + 4, 0); std::vector<int> st(F + 8);
        if (traverse_table(tv[t], lmc.data(), F, fv.data(), v2d1[t].data(), d2c[t].data(), st.data(), nvt, &entries[t])) return 0;
    }
    // ---- quantise
    float pmin[3] = {1e30f, 1e30f, 1e30f}, pmax[3] = {-1e30f, -1e30f, -1e30f};
    for (uint32_t v = 0; v < nv; v++) for (int k = 0; k < 3; k++) { pmin[k] = std::min(pmin[k], pos[3 * v + k]); pmax[k] = std::max(pmax[k], pos[3 * v + k]); }
    float prange = std::max(pmax[0] - pmin[0], std::max(pmax[1] - pmin[1], pmax[2] - pmin[2])); if (prange <= 0) prange = 1;
    const int32_t pmaxq = (1 << qp) - 1;
    auto quant = [](float v, float mn, float range, int32_t maxq) { double q = floor(((double)v - mn) / range * maxq + 0.5); if (q < 0) q = 0; if (q > maxq) q = maxq; return (int32_t)q; };
    float umin[2] = {1e30f, 1e30f}, umax[2] = {-1e30f, -1e30f};
    for (uint32_t v = 0; v < nuv; v++) for (int k = 0; k < 2; k++) { umin[k] = std::min(umin[k], uv[2 * v + k]); umax[k] = std::max(umax[k], uv[2 * v + k]); }
    float urange = std::max(umax[0] - umin[0], umax[1] - umin[1]); if (urange <= 0) urange = 1;
    const int32_t umaxq = (1 << qt) - 1;
    // ---- POSITION: parallelogram + wrap in entry order
    const int np = (int)entries[0];
    std::vector<int32_t> posq((size_t)np * 3); std::vector<uint32_t> possym((size_t)np * 3);
    for (int p = 0; p < np; p++) { const uint32_t ov = faces_pos[d2o[d2c[0][p]]]; for (int k = 0; k < 3; k++) posq[3 * p + k] = quant(pos[3 * ov + k], pmin[k], prange, pmaxq); }
    int32_t pwmin = posq[0], pwmax = posq[0]; for (int32_t q : posq) { pwmin = std::min(pwmin, q); pwmax = std::max(pwmax, q); }
    for (int p = 0; p < np; p++) {
        int par[4] = {-1, -1, -1, 0};
        if (p > 0) parallelogram_parents(p, tv[0], d2c[0].data(), v2d1[0].data(), par);
        for (int k = 0; k < 3; k++) {
            long long pred = 0;
            if (p > 0) pred = par[0] >= 0 ? ((long long)posq[3 * par[1] + k] + posq[3 * par[2] + k]) - posq[3 * par[0] + k] : posq[3 * (p - 1) + k];
            possym[3 * p + k] = zigzag_sym(wrap_corr(posq[3 * p + k], pred, pwmin, pwmax));
        }
    }
    // ---- TEX_COORD: portable predictor, orientation chosen by the encoder
    const int nu = (int)entries[1];
    std::vector<int32_t> uvq((size_t)nu * 2); std::vector<uint32_t> uvsym((size_t)nu * 2); std::vector<uint8_t> orient_enc;
    for (int p = 0; p < nu; p++) { const uint32_t ou = faces_uv[d2o[d2c[1][p]]]; for (int k = 0; k < 2; k++) uvq[2 * p + k] = quant(uv[2 * ou + k], umin[k], urange, umaxq); }
    int32_t uwmin = uvq[0], uwmax = uvq[0]; for (int32_t q : uvq) { uwmin = std::min(uwmin, q); uwmax = std::max(uwmax, q); }
    for (int p = 0; p < nu; p++) {
        UvPrep q; uv_prepare(p, tv[1], d2c[1].data(), v2d1[1].data(), v2d1[0].data(), posq.data(), q);
        long long pred[2]; bool have = false;
        if (q.pd < p && q.nd < p && q.pd >= 0 && q.nd >= 0) {
            const long long nuv[2] = {uvq[q.nd * 2], uvq[q.nd * 2 + 1]}, puv[2] = {uvq[q.pd * 2], uvq[q.pd * 2 + 1]};
            if (nuv[0] == puv[0] && nuv[1] == puv[1]) { pred[0] = puv[0]; pred[1] = puv[1]; have = true; }
            else if (q.pn2 != 0) {
                const long long pnuv[2] = {puv[0] - nuv[0], puv[1] - nuv[1]};
                const long long xuv[2] = {nuv[0] * q.pn2 + q.dot * pnuv[0], nuv[1] * q.pn2 + q.dot * pnuv[1]};
                const long long cxuv[2] = {pnuv[1] * q.ns, -pnuv[0] * q.ns};
                long long a[2] = {(int32_t)((xuv[0] + cxuv[0]) / q.pn2), (int32_t)((xuv[1] + cxuv[1]) / q.pn2)};
                long long b[2] = {(int32_t)((xuv[0] - cxuv[0]) / q.pn2), (int32_t)((xuv[1] - cxuv[1]) / q.pn2)};
                const long long da = (a[0] - uvq[2 * p]) * (a[0] - uvq[2 * p]) + (a[1] - uvq[2 * p + 1]) * (a[1] - uvq[2 * p + 1]);
                const long long db = (b[0] - uvq[2 * p]) * (b[0] - uvq[2 * p]) + (b[1] - uvq[2 * p + 1]) * (b[1] - uvq[2 * p + 1]);
                const bool o = da <= db; orient_enc.push_back(o);
                pred[0] = o ? a[0] : b[0]; pred[1] = o ? a[1] : b[1]; have = true;
            }
        }
        if (!have) {
            if (q.nd < p && q.nd >= 0) { pred[0] = uvq[q.nd * 2]; pred[1] = uvq[q.nd * 2 + 1]; }
            else if (p > 0) { pred[0] = uvq[(p - 1) * 2]; pred[1] = uvq[(p - 1) * 2 + 1]; }
            else pred[0] = pred[1] = 0;
        }
        for (int k = 0; k < 2; k++) uvsym[2 * p + k] = zigzag_sym(wrap_corr(uvq[2 * p + k], pred[k], uwmin, uwmax));
    }
    // the decoder pops orientations from the BACK of the decoded list, and decodes "same as previous" bits
    std::vector<uint8_t> orient_bits(orient_enc.size());
    { const size_t k = orient_enc.size(); int last = 1; for (size_t i = 0; i < k; i++) { const int o = orient_enc[k - 1 - i]; orient_bits[i] = (o == last); last = o; } }
    // ---- NORMAL: geometric prediction + canonicalised octahedron, positive corrections
    const int nn = (int)entries[2];
    const int32_t MAXQ = (1 << qn) - 1, MAXV = MAXQ - 1, CEN = MAXV / 2;
    std::vector<uint32_t> nrmsym((size_t)nn * 2); std::vector<uint8_t> flipbits(nn);
    for (int p = 0; p < nn; p++) {
        const uint32_t ov = faces_pos[d2o[d2c[2][p]]]; const float *nv3 = nrm + 3 * ov;
        // FloatVectorToQuantizedOctahedralCoords
        const double as = fabs(nv3[0]) + fabs(nv3[1]) + fabs(nv3[2]); double sc[3] = {1, 0, 0};
        if (as > 1e-6) { sc[0] = nv3[0] / as; sc[1] = nv3[1] / as; sc[2] = nv3[2] / as; }
        int32_t iv[3]; iv[0] = (int32_t)floor(sc[0] * CEN + 0.5); iv[1] = (int32_t)floor(sc[1] * CEN + 0.5); iv[2] = CEN - abs(iv[0]) - abs(iv[1]);
        if (iv[2] < 0) { if (iv[1] > 0) iv[1] += iv[2]; else iv[1] -= iv[2]; iv[2] = 0; }
        if (sc[2] < 0) iv[2] = -iv[2];
        int32_t s, t;
        if (iv[0] >= 0) { s = iv[1] + CEN; t = iv[2] + CEN; } else { s = iv[1] < 0 ? abs(iv[2]) : MAXV - abs(iv[2]); t = iv[2] < 0 ? abs(iv[1]) : MAXV - abs(iv[1]); }
        if ((s == 0 && t == 0) || (s == 0 && t == MAXV) || (s == MAXV && t == 0)) { s = MAXV; t = MAXV; }
        else if (s == 0 && t > CEN) t = CEN - (t - CEN); else if (s == MAXV && t < CEN) t = CEN + (CEN - t);
        else if (t == MAXV && s < CEN) s = CEN + (CEN - s); else if (t == 0 && s > CEN) s = CEN - (s - CEN);
        int best = -1; int32_t bc[2] = {0, 0}; long long bcost = 0;
        for (int flip = 0; flip < 2; flip++) {
            int32_t ps, pt; normal_predict_oct(p, tv[2], d2c[2].data(), v2d1[0].data(), posq.data(), flip, MAXQ, &ps, &pt);
            // forward transform = inverse of oct_apply_correction
            int32_t px = ps - CEN, py = pt - CEN, ox = s - CEN, oy = t - CEN;
            const int ind = abs(px) + abs(py) <= CEN;
            if (!ind) { oct_invert_diamond(px, py, CEN); oct_invert_diamond(ox, oy, CEN); }
            const int bl = (px == 0 && py == 0) || (px < 0 && py <= 0);
            int rc; if (px == 0) rc = py == 0 ? 0 : (py > 0 ? 3 : 1); else if (px > 0) rc = py >= 0 ? 2 : 1; else rc = py <= 0 ? 0 : 3;
            if (!bl) { oct_rotate(px, py, rc); oct_rotate(ox, oy, rc); }
            int32_t c0 = ox - px, c1 = oy - py; if (c0 < 0) c0 += MAXQ; if (c1 < 0) c1 += MAXQ;
            int32_t vs, vt; oct_apply_correction(ps, pt, c0, c1, MAXQ, &vs, &vt);
            if (vs != s || vt != t) continue;       // this orientation cannot reproduce the value exactly
            const long long cost = std::min(c0, MAXQ - c0) + std::min(c1, MAXQ - c1);
            if (best < 0 || cost < bcost) { best = flip; bc[0] = c0; bc[1] = c1; bcost = cost; }
        }
        if (best < 0) return 0;
        flipbits[p] = (uint8_t)best; nrmsym[2 * p] = (uint32_t)bc[0]; nrmsym[2 * p + 1] = (uint32_t)bc[1];
    }
    // ---- assemble
    Bytes b; b.reserve((size_t)nv * 6);
    b.insert(b.end(), {'D', 'R', 'A', 'C', 'O'}); put_u8(b, 2); put_u8(b, 2); put_u8(b, 1); put_u8(b, 1); put_u16(b, 0);
    put_u8(b, 2); put_varint(b, nv); put_varint(b, nf); put_u8(b, 2); put_varint(b, (uint64_t)nsym); put_varint(b, (uint64_t)nsplit);
    put_varint(b, 0);                                         // no topology split events (genus 0, closed)
    put_bytes(b, sf_run);
    for (int i = 0; i < 2; i++) write_rabs(b, seambits[i]);
    for (int i = 0; i < 6; i++) { put_varint(b, ctxs[i].size()); if (!ctxs[i].empty()) write_symbols_raw(b, ctxs[i]); }
    put_u8(b, 3);                                             // attribute decoders
    put_u8(b, 0xFF); put_u8(b, 0); put_u8(b, 0);              // {att_data_id=-1, MESH_VERTEX_ATTRIBUTE, DEPTH_FIRST}
    put_u8(b, 0); put_u8(b, 1); put_u8(b, 0);                 // {0, MESH_CORNER_ATTRIBUTE, DEPTH_FIRST}
    put_u8(b, 1); put_u8(b, 1); put_u8(b, 0);                 // {1, MESH_CORNER_ATTRIBUTE, DEPTH_FIRST}
    put_varint(b, 1); put_u8(b, 0); put_u8(b, 9); put_u8(b, 3); put_u8(b, 0); put_varint(b, 0); put_u8(b, 2);   // POSITION f32x3, QUANTIZATION
    put_varint(b, 1); put_u8(b, 3); put_u8(b, 9); put_u8(b, 2); put_u8(b, 0); put_varint(b, 1); put_u8(b, 2);   // TEX_COORD f32x2, QUANTIZATION
    put_varint(b, 1); put_u8(b, 1); put_u8(b, 9); put_u8(b, 3); put_u8(b, 0); put_varint(b, 2); put_u8(b, 3);   // NORMAL f32x3, NORMALS
    // decoder 0 portable data + transform data
    put_u8(b, 1); put_u8(b, 1); put_u8(b, 1); write_symbols(b, possym, 3, 0); put_u32(b, (uint32_t)pwmin); put_u32(b, (uint32_t)pwmax);
    for (int k = 0; k < 3; k++) put_f32(b, pmin[k]);
    put_f32(b, prange); put_u8(b, (uint32_t)qp);
    // decoder 1
    put_u8(b, 5); put_u8(b, 1); put_u8(b, 1); write_symbols(b, uvsym, 2, 1);
    put_u32(b, (uint32_t)orient_bits.size()); write_rabs(b, orient_bits); put_u32(b, (uint32_t)uwmin); put_u32(b, (uint32_t)uwmax);
    for (int k = 0; k < 2; k++) put_f32(b, umin[k]);
    put_f32(b, urange); put_u8(b, (uint32_t)qt);
    // decoder 2
    put_u8(b, 6); put_u8(b, 3); put_u8(b, 1); write_symbols(b, nrmsym, 2, 2);
    put_u32(b, (uint32_t)MAXQ); put_u32(b, (uint32_t)CEN); write_rabs(b, flipbits);
    put_u8(b, (uint32_t)qn);
    *out_buf = (uint8_t *)malloc(b.size() + 16);
    memcpy(*out_buf, b.data(), b.size());
    return b.size();
}

extern "C" void uvsynth_free(void *p) { free(p); }
extern "C" void uvsynth_draco_tagged(int mask) { g_tagged_mask = mask; }
