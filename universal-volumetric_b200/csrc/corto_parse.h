// corto_parse.h -- HOST structural parse of a Corto .crt frame into a CortoFrame descriptor (header + section walk:
// deprecated/encoder/dev/src/decoder.cpp:41-85, index_attribute.h:83-98, cstream.h:285-362).  Only sizes and offsets are read; every
// payload byte is decoded on the GPU (corto_decode.cu).  Kept in a header so that the guard-page fuzzer (tests/tools/fuzz_host.cpp)
// runs exactly the code the library runs on untrusted bytes.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "uvol_internal.h"

namespace {

struct TunBlock { uint32_t probs_off, nsym, size, csize, data_off; uint64_t o_out; };      // offsets inside the file; o_out in scratch
struct BitBlock { uint32_t nwords, data_off; };
enum { CK_POSITION = 0, CK_UV = 1, CK_NORMAL = 2, CK_COLOR = 3 };
struct CortoAttr {
    int32_t kind, N /*header components*/, vN /*values per vertex in the stream*/, strategy, nlogs, pred /*normal prediction: 0 DIFF 1 ESTIMATED 2 BORDER*/;
    float q; uint32_t qc[4]; uint32_t count /*values in the stream*/;
    BitBlock bits; TunBlock logs[4]; uint64_t o_val, out;
};
#define CORTO_MAX_ATTRS 4
struct CortoFrame {
    uint64_t file_off; uint32_t file_len; int32_t status;
    uint32_t nvert, nface, ngroups, groups_off /*aux u32: end face per group*/, max_front, entropy;
    TunBlock clers; BitBlock ibits;
    int32_t nattr, pos_attr, nrm_attr, index16; CortoAttr attr[CORTO_MAX_ATTRS];
    uint64_t o_front, o_third, o_queue, o_delayed, o_pred, out_index, out_index16;
    uint64_t o_voff, o_vfill, o_vlist, o_est, o_bflag;          // normal estimation: corner lists per vertex, summed face normals, boundary marks / their scan
};
struct CJob { uint32_t frame; int32_t what; };

struct Rd {
    const uint8_t *b; size_t n, p; bool err;
    uint8_t u8() { if (p + 1 > n) { err = true; return 0; } return b[p++]; }
    uint16_t u16() { uint16_t a = u8(), c = u8(); return (uint16_t)(a | (c << 8)); }
    uint32_t u32() { if (p + 4 > n) { err = true; p = n; return 0; } uint32_t v; memcpy(&v, b + p, 4); p += 4; return v; }
    float f32() { uint32_t v = u32(); float f; memcpy(&f, &v, 4); return f; }
    std::string str() { uint16_t l = u16(); if (err || l > n - p) { err = true; return ""; } std::string s((const char *)b + p, l ? l - 1 : 0); p += l; return s; }
};
bool read_tunstall(Rd &r, uint32_t entropy, TunBlock &t) {
    memset(&t, 0, sizeof t);
    if (entropy == 0) { t.nsym = 0xffffffffu; t.size = r.u32(); t.csize = t.size; t.data_off = (uint32_t)r.p; if (r.err || t.size > r.n - r.p) return false; r.p += t.size; return true; }
    t.nsym = r.u8(); t.probs_off = (uint32_t)r.p;
    if (r.err || 2 * (size_t)t.nsym > r.n - r.p) return false;
    r.p += 2 * (size_t)t.nsym;
    t.size = r.u32(); t.csize = r.u32(); t.data_off = (uint32_t)r.p;
    // a code byte expands to at most one dictionary word; the decoded size a header may claim is bounded by what its bytes can
    // encode (words of up to 2 KiB) and by 2^28 -- checked before anything is reserved
    if (r.err || t.csize > r.n - r.p || t.size > (1u << 28) || (uint64_t)t.size > 2048ull * t.csize + 64) return false;
    r.p += t.csize;
    return true;
}
bool read_bits(Rd &r, BitBlock &b) {
    b.nwords = r.u32();
    const size_t pad = r.p & 3; if (pad) r.p += 4 - pad;
    b.data_off = (uint32_t)r.p;
    if (r.err || r.p > r.n || 4ull * b.nwords > r.n - r.p) return false;
    r.p += 4ull * b.nwords;
    return true;
}

// Header + section walk (decoder.cpp:41-85, index_attribute.h:83-98, cstream.h:285-362).
int corto_parse(const uint8_t *data, size_t len, CortoFrame &f, std::vector<uint32_t> &aux, uint64_t max_faces) {
    Rd r{data, len, 0, false};
    if (len < 24 || r.u32() != 0x787A6300u) return UVOL_ERR_CORRUPT;
    (void)r.u32();
    f.entropy = r.u8();
    if (f.entropy > 1) return UVOL_ERR_UNSUPPORTED;
    const uint32_t nexif = r.u32();
    if (r.err || nexif > 4096) return UVOL_ERR_CORRUPT;
    for (uint32_t i = 0; i < nexif; i++) { r.str(); r.str(); }
    const uint32_t nattr = r.u32();
    if (r.err || nattr > 16) return UVOL_ERR_CORRUPT;
    struct Hdr { std::string name; int codec; float q; int N, format, strategy; };
    std::vector<Hdr> hdr(nattr);
    for (auto &h : hdr) { h.name = r.str(); h.codec = (int)r.u32(); h.q = r.f32(); h.N = r.u8(); h.format = r.u8(); h.strategy = r.u8(); }
    f.nvert = r.u32(); f.nface = r.u32();
    if (r.err || f.nvert == 0 || f.nvert > (1u << 26) || f.nface > (1u << 27)) return UVOL_ERR_CORRUPT;
    if (f.nface == 0) return UVOL_ERR_UNSUPPORTED;                 // point clouds: DecodeMesh returns -1 (corto_codec.cpp:27-30)
    if ((uint64_t)f.nface > max_faces) return UVOL_ERR_UNSUPPORTED; // resource limit (uvol_config.max_faces_per_frame)
    f.ngroups = r.u32(); f.groups_off = (uint32_t)aux.size();
    if (r.err || f.ngroups > 65536) return UVOL_ERR_CORRUPT;
    for (uint32_t g = 0; g < f.ngroups; g++) {
        const uint32_t end = r.u32(); const uint8_t np = r.u8();
        for (int k = 0; k < np; k++) { r.str(); r.str(); }
        if (r.err || end > f.nface) return UVOL_ERR_CORRUPT;
        aux.push_back(end);
    }
    f.max_front = r.u32();
    if (!read_tunstall(r, f.entropy, f.clers) || !read_bits(r, f.ibits)) return UVOL_ERR_TRUNCATED;
    // every face and every vertex costs at least one connectivity symbol: counts a header claims beyond that are rejected here,
    // per item, before any arena is planned from them
    if ((uint64_t)f.nface > (uint64_t)f.clers.size + 1 || (uint64_t)f.nvert > 3ull * f.clers.size + 3) return UVOL_ERR_CORRUPT;
    // attributes follow in std::map (alphabetical) order of their names (decoder.cpp:146-147)
    std::vector<int> order(nattr); for (uint32_t i = 0; i < nattr; i++) order[i] = (int)i;
    for (uint32_t i = 0; i < nattr; i++) for (uint32_t j = i + 1; j < nattr; j++) if (hdr[order[j]].name < hdr[order[i]].name) std::swap(order[i], order[j]);
    f.nattr = 0; f.pos_attr = f.nrm_attr = -1;
    for (uint32_t i = 0; i < nattr; i++) {
        const Hdr &h = hdr[order[i]];
        if (h.N < 1 || h.N > 4) return UVOL_ERR_UNSUPPORTED;
        CortoAttr a; memset(&a, 0, sizeof a);
        a.N = h.N; a.strategy = h.strategy; a.q = h.q; a.kind = -1;
        if (h.codec == 2) {                                   // NormalAttr::decode (normal_attribute.cpp:168-175): prediction byte, then a correlated array of 2
            a.kind = h.name == "normal" ? CK_NORMAL : -1; a.vN = 2; a.nlogs = 1;
            a.pred = r.u8();
            if (r.err || a.pred > 2) return UVOL_ERR_CORRUPT;
        } else if (h.codec == 3) {                            // ColorAttr::decode (color_attribute.h:55-59): one step byte per component, then per-component values
            a.kind = h.name == "color" ? CK_COLOR : -1; a.vN = h.N; a.nlogs = h.N;
            for (int k = 0; k < 4; k++) a.qc[k] = k < 3 ? 4 : 8;
            for (int k = 0; k < h.N; k++) a.qc[k] = r.u8();
        } else {
            a.kind = h.name == "position" ? CK_POSITION : (h.name == "uv" ? CK_UV : -1); a.vN = h.N; a.nlogs = (h.strategy & 2) ? 1 : h.N;
            if ((a.kind == CK_POSITION && a.N != 3) || (a.kind == CK_UV && a.N != 2)) return UVOL_ERR_UNSUPPORTED;
        }
        if (!read_bits(r, a.bits)) return UVOL_ERR_TRUNCATED;
        for (int k = 0; k < a.nlogs; k++) {
            if (!read_tunstall(r, f.entropy, a.logs[k])) return UVOL_ERR_TRUNCATED;
            if (a.kind == CK_NORMAL ? a.logs[k].size > f.nvert : a.logs[k].size != f.nvert) return UVOL_ERR_CORRUPT;
        }
        a.count = a.logs[0].size;
        if (a.kind == CK_NORMAL && a.pred != 2 && a.count != f.nvert) return UVOL_ERR_CORRUPT;
        if (a.kind < 0) continue;                             // not exported: parsed past
        for (int k = 0; k < f.nattr; k++) if (f.attr[k].kind == a.kind) return UVOL_ERR_CORRUPT;
        if (f.nattr >= CORTO_MAX_ATTRS) return UVOL_ERR_UNSUPPORTED;
        if (a.kind == CK_POSITION) f.pos_attr = f.nattr;
        if (a.kind == CK_NORMAL) f.nrm_attr = f.nattr;
        f.attr[f.nattr++] = a;
    }
    if (f.pos_attr < 0) return UVOL_ERR_UNSUPPORTED;
    return UVOL_OK;
}

}  // namespace
