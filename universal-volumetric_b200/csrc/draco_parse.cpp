// draco_parse.cpp -- host-side structural parse of a .drc file into a DracoFrame descriptor.
//
// Replaces the header walk that draco::Decoder::DecodeArrayToMesh performs before any arithmetic
// (reference call site src/lib/DRACOLoader.js:478-483).  It touches only section headers, varints
// and probability tables -- a few hundred bytes per frame; every payload byte (rANS / rABS runs,
// raw bit fields) is decoded on the GPU.  Bitstream layout: SURVEY.md Appendix A.2.
#include <string.h>
#include <vector>
#include "uvol_internal.h"

namespace {
struct Rd {
    const uint8_t *b; size_t n, p; bool err;
    uint8_t u8() { if (p + 1 > n) { err = true; return 0; } return b[p++]; }
    uint16_t u16() { uint16_t a = u8(), c = u8(); return (uint16_t)(a | (c << 8)); }
    uint32_t u32() { if (p + 4 > n) { err = true; p = n; return 0; } uint32_t v; memcpy(&v, b + p, 4); p += 4; return v; }
    float f32() { uint32_t v = u32(); float f; memcpy(&f, &v, 4); return f; }
    uint64_t varint() {
        uint64_t v = 0; int sh = 0;
        for (;;) { uint8_t c = u8(); if (err) return 0; v |= (uint64_t)(c & 0x7f) << sh; sh += 7; if (!(c & 0x80)) return v; if (sh > 63) { err = true; return 0; } }
    }
};

bool read_rabs(Rd &r, RabsStream &s) {
    s.prob_zero = r.u8(); uint64_t sz = r.varint();
    if (r.err || sz == 0 || sz > r.n - r.p) return false;
    s.data_off = (uint32_t)r.p; s.data_len = (uint32_t)sz; s.pad = 0; r.p += sz;
    return true;
}

// Metadata section (header flag 0x8000, draco::MetadataEncoder::EncodeGeometryMetadata): walked past, never used -- the reference's
// loader reads no metadata (src/lib/DRACOLoader.js:470-554).  metadata = varint entries {u8 name length, name, varint value size, value},
// varint sub-metadata count {u8 name length, name, metadata}.
bool skip_metadata(Rd &r, int depth) {
    if (depth > 32) return false;
    const uint64_t ne = r.varint();
    if (r.err || ne > r.n) return false;
    for (uint64_t i = 0; i < ne; i++) {
        const uint8_t nl = r.u8(); if (r.err || nl > r.n - r.p) return false; r.p += nl;
        const uint64_t vs = r.varint(); if (r.err || vs > r.n - r.p) return false; r.p += vs;
    }
    const uint64_t ns = r.varint();
    if (r.err || ns > r.n) return false;
    for (uint64_t i = 0; i < ns; i++) {
        const uint8_t nl = r.u8(); if (r.err || nl > r.n - r.p) return false; r.p += nl;
        if (!skip_metadata(r, depth + 1)) return false;
    }
    return true;
}

// Probability table of a rANS symbol run (alphabet size varint, then one token per symbol or zero run), appended to `aux`.
int read_prob_table(Rd &r, RansStream &s, std::vector<uint32_t> &aux) {
    uint64_t n = r.varint();
    if (r.err || n == 0 || n > (1u << 20)) return UVOL_ERR_CORRUPT;
    s.alphabet = (uint32_t)n; s.prob_off = (uint32_t)aux.size();
    aux.resize(aux.size() + n, 0u);
    uint32_t *prob = aux.data() + s.prob_off;
    for (uint32_t i = 0; i < n;) {
        uint8_t d = r.u8(); unsigned tok = d & 3;
        if (r.err) return UVOL_ERR_TRUNCATED;
        if (tok == 3) { uint32_t run = (d >> 2) + 1; if (i + run > n) return UVOL_ERR_CORRUPT; i += run; }
        else { uint32_t pr = d >> 2; for (unsigned k = 0; k < tok; k++) pr |= (uint32_t)r.u8() << (8 * (k + 1) - 2); prob[i++] = pr; }
    }
    s.nnz = 0; for (uint32_t i = 0; i < n; i++) s.nnz += aux[s.prob_off + i] != 0;
    return UVOL_OK;
}

// The TAGGED scheme stores no length for its raw bit fields: they end where the tags say.  To find the data that follows (the
// prediction parameters of this attribute, every later attribute), the parser walks the tag run once -- to the coder's terminal
// state, the tag count is not in the file either -- and adds up the bit lengths.  This locates bytes; the tags the decode uses are
// decoded again on the device with everything else.  false: not a valid run (or more tags than the mesh has corners).
bool sum_tags(const uint8_t *data, uint32_t nbytes, const uint32_t *prob, uint32_t alphabet, uint32_t pb, uint64_t max_tags, uint64_t *ntags, uint64_t *sum) {
    *ntags = 0; *sum = 0;
    const uint32_t prec = 1u << pb, lbase = prec * 4u;
    if (alphabet > 64 || nbytes == 0) return false;                 // a tag is a bit length: 0..32
    uint32_t cum[65]; cum[0] = 0;
    for (uint32_t i = 0; i < alphabet; i++) { if (prob[i] > prec) return false; cum[i + 1] = cum[i] + prob[i]; }
    if (cum[alphabet] != prec) return false;
    const unsigned k = (data[nbytes - 1] >> 6) + 1u;
    if (nbytes < k) return false;
    uint32_t st = 0; for (unsigned i = 0; i < k; i++) st |= (uint32_t)data[nbytes - k + i] << (8 * i);
    st = (st & ((1u << (8 * k - 2)) - 1u)) + lbase;
    uint32_t off = nbytes - k;
    for (;;) {
        while (st < lbase && off > 0) st = st * 256u + data[--off];
        if (off == 0 && st == lbase) return true;
        if (*ntags >= max_tags) return false;
        const uint32_t q = st >> pb, rem = st & (prec - 1);
        uint32_t sy = 0; while (cum[sy + 1] <= rem) sy++;
        st = q * prob[sy] + rem - cum[sy];
        if (sy > 32) return false;
        *sum += sy; ++*ntags;
    }
}

// DecodeSymbols header: scheme, then RAW (max_bit_length, probability table, byte run) or TAGGED (probability table of the tags, byte
// run, raw bit fields: located by walking the tags, see above).  nc = components per value tuple; tagged = null: RAW only.
int read_symbols(Rd &r, RansStream &s, std::vector<uint32_t> &aux, int nc = 1, uint64_t max_tuples = 0, DracoAttr *tagged = nullptr) {
    uint8_t scheme = r.u8();
    if (r.err) return UVOL_ERR_TRUNCATED;
    if (scheme > 1 || (scheme == 0 && !tagged)) return UVOL_ERR_UNSUPPORTED;
    int pb = 12;
    if (scheme == 1) { const uint8_t mbl = r.u8(); pb = (3 * mbl) / 2; if (pb < 12) pb = 12; if (pb > 20) pb = 20; }
    int rc = read_prob_table(r, s, aux); if (rc) return rc;
    s.pb = (uint32_t)pb;
    uint64_t nb = r.varint();
    if (r.err || nb > r.n - r.p) return UVOL_ERR_TRUNCATED;
    s.data_off = (uint32_t)r.p; s.data_len = (uint32_t)nb; r.p += nb;
    if (scheme == 0) {
        uint64_t ntags, sum;
        if (!sum_tags(r.b + s.data_off, s.data_len, aux.data() + s.prob_off, s.alphabet, s.pb, max_tuples, &ntags, &sum)) return UVOL_ERR_CORRUPT;
        const uint64_t bytes = (sum * (uint64_t)nc + 7) >> 3;
        if (bytes > r.n - r.p) return UVOL_ERR_TRUNCATED;
        tagged->tagged = 1; tagged->tag_bits_off = (uint32_t)r.p; tagged->tag_bits_len = (uint32_t)bytes; r.p += bytes;
    }
    return UVOL_OK;
}
}  // namespace

// Parses `data` into `f` (file_off/file_len are set by the caller).  Probability tables and
// topology-split events are appended to `aux`.  Returns UVOL_OK or a negative status.
int uvol_draco_parse(const uint8_t *data, size_t len, DracoFrame &f, std::vector<uint32_t> &aux) {
    Rd r{data, len, 0, false};
    if (len < 11 || memcmp(data, "DRACO", 5)) return UVOL_ERR_CORRUPT;
    r.p = 5;
    int maj = r.u8(), mino = r.u8(), etype = r.u8(), meth = r.u8(); int flags = r.u16();
    if (maj != 2 || mino != 2 || etype != 1 || meth != 1) return UVOL_ERR_UNSUPPORTED;
    if (flags & 0x8000) {          // geometry metadata: {varint attribute count, each: varint unique id + metadata}, then the geometry's own
        const uint64_t na = r.varint();
        if (r.err || na > len) return UVOL_ERR_CORRUPT;
        for (uint64_t i = 0; i < na; i++) { (void)r.varint(); if (!skip_metadata(r, 0)) return UVOL_ERR_CORRUPT; }
        if (!skip_metadata(r, 0)) return UVOL_ERR_CORRUPT;
        // every offset recorded below is relative to the start of the file: nothing else changes
    }
    f.trav = r.u8(); f.nv_enc = (uint32_t)r.varint(); f.nf = (uint32_t)r.varint(); f.nad = r.u8();
    f.nsym = (uint32_t)r.varint(); f.nsplit = (uint32_t)r.varint();
    if (r.err) return UVOL_ERR_TRUNCATED;
    if (f.trav != 0 && f.trav != 2) return UVOL_ERR_UNSUPPORTED;
    if (f.nad > UVOL_MAX_ATTR_DATA) return UVOL_ERR_UNSUPPORTED;
    if (f.nf == 0 || f.nf > (1u << 26) || f.nsym > f.nf || f.nv_enc > 3 * f.nf + 3 || f.nsplit > f.nf) return UVOL_ERR_CORRUPT;
    f.nts = (uint32_t)r.varint();
    if (r.err || f.nts > f.nf) return UVOL_ERR_CORRUPT;
    f.ts_off = (uint32_t)aux.size();
    aux.resize(aux.size() + 3 * (size_t)f.nts);
    uint32_t last = 0;
    for (uint32_t i = 0; i < f.nts; i++) {
        uint32_t d = (uint32_t)r.varint(); uint32_t src = last + d; uint32_t d2 = (uint32_t)r.varint();
        aux[f.ts_off + 3 * i] = src; aux[f.ts_off + 3 * i + 1] = src - d2; last = src;
    }
    if (f.nts) {
        size_t nb = ((size_t)f.nts + 7) / 8;
        if (r.p + nb > r.n) return UVOL_ERR_TRUNCATED;
        for (uint32_t i = 0; i < f.nts; i++) aux[f.ts_off + 3 * i + 2] = (data[r.p + (i >> 3)] >> (i & 7)) & 1;
        r.p += nb;
    }
    f.stdsym_off = f.stdsym_len = 0;
    if (f.trav == 0) {
        uint64_t sz = r.varint();
        if (r.err || sz > r.n - r.p) return UVOL_ERR_TRUNCATED;
        f.stdsym_off = (uint32_t)r.p; f.stdsym_len = (uint32_t)sz; r.p += sz;
    }
    if (!read_rabs(r, f.start_faces)) return UVOL_ERR_TRUNCATED;
    for (uint32_t i = 0; i < f.nad; i++) if (!read_rabs(r, f.seams[i])) return UVOL_ERR_TRUNCATED;
    for (int i = 0; i < 6; i++) memset(&f.ctx[i], 0, sizeof(RansStream));
    if (f.trav == 2) {
        uint64_t total = 0;
        for (int i = 0; i < 6; i++) {
            uint64_t n = r.varint();
            if (r.err || n > f.nsym) return UVOL_ERR_CORRUPT;
            if (n) { int rc = read_symbols(r, f.ctx[i], aux); if (rc) return rc; }
            f.ctx[i].count = (uint32_t)n; total += n;
        }
        if (f.nsym && total != (uint64_t)f.nsym - 1) return UVOL_ERR_CORRUPT;
    }
    // ---- attribute decoders
    int ndec = r.u8();
    if (r.err || ndec < 1 || ndec > 8) return UVOL_ERR_CORRUPT;
    struct Dec { int att_data_id, dec_type, trav, natt, first; } dec[8];
    for (int i = 0; i < ndec; i++) { dec[i].att_data_id = (int8_t)r.u8(); dec[i].dec_type = r.u8(); dec[i].trav = r.u8(); }
    f.nattr = 0; f.pos_attr = -1;
    for (int i = 0; i < ndec; i++) {
        Dec &d = dec[i];
        if (d.att_data_id >= (int)f.nad || d.trav != 0 || d.dec_type > 1) return UVOL_ERR_UNSUPPORTED;
        if (d.dec_type == 1 && d.att_data_id < 0) return UVOL_ERR_CORRUPT;
        // the count is attacker-controlled: compare as unsigned 64-bit BEFORE any int arithmetic (a varint of INT_MAX used to wrap
        // the signed sum and let the loop below write past f.attr[])
        const uint64_t na = r.varint(); d.first = f.nattr;
        if (r.err || na < 1 || na > (uint64_t)(UVOL_MAX_ATTRS - f.nattr)) return UVOL_ERR_UNSUPPORTED;
        d.natt = (int)na;
        for (int j = 0; j < d.natt; j++) {
            DracoAttr &a = f.attr[f.nattr + j]; memset(&a, 0, sizeof a);
            a.type = (int8_t)r.u8(); a.dtype = (int8_t)r.u8(); a.nc = (int8_t)r.u8(); a.normalized = (int8_t)r.u8(); (void)r.varint();
            if (a.nc < 1 || a.nc > 4) return UVOL_ERR_UNSUPPORTED;
            a.table = (int8_t)(d.dec_type == 1 ? d.att_data_id : -1);
            a.out_slot = -1;
        }
        for (int j = 0; j < d.natt; j++) f.attr[f.nattr + j].seq = (int8_t)r.u8();
        f.nattr += d.natt;
    }
    if (r.err) return UVOL_ERR_TRUNCATED;
    bool sem_done[4] = {false, false, false, false};
    for (int i = 0; i < ndec; i++) {
        Dec &d = dec[i];
        for (int j = 0; j < d.natt; j++) {
            DracoAttr &a = f.attr[d.first + j];
            if (a.seq < 1 || a.seq > 3) return UVOL_ERR_UNSUPPORTED;
            a.vnc = a.seq == 3 ? 2 : a.nc;
            a.pred = (int8_t)r.u8(); a.xform = -1;
            if (a.pred != -2) a.xform = (int8_t)r.u8();
            int compressed = r.u8();
            if (r.err) return UVOL_ERR_TRUNCATED;
            if (!compressed) return UVOL_ERR_UNSUPPORTED;     // raw ints need the entry count to be skipped
            int rc = read_symbols(r, a.sym, aux, a.vnc, 3ull * f.nf + 8, &a); if (rc) return rc;
            a.sym.count = 0xFFFFFFFFu;
            if (a.pred == -2) { /* no prediction data */ }
            else if (a.pred == 0 || a.pred == 1) {
                if (a.xform != 1) return UVOL_ERR_UNSUPPORTED;
                a.wmin = (int32_t)r.u32(); a.wmax = (int32_t)r.u32();
                if (a.wmax < a.wmin) return UVOL_ERR_CORRUPT;
            } else if (a.pred == 5) {
                if (a.xform != 1 || a.nc != 2 || f.pos_attr < 0) return UVOL_ERR_UNSUPPORTED;
                a.num_orient = (int32_t)r.u32();
                if (r.err || a.num_orient < 0) return UVOL_ERR_CORRUPT;
                if (!read_rabs(r, a.aux_bits)) return UVOL_ERR_TRUNCATED;
                a.wmin = (int32_t)r.u32(); a.wmax = (int32_t)r.u32();
                if (a.wmax < a.wmin) return UVOL_ERR_CORRUPT;
            } else if (a.pred == 6) {
                if (a.xform != 3 || a.seq != 3 || f.pos_attr < 0) return UVOL_ERR_UNSUPPORTED;
                a.wmin = (int32_t)r.u32(); a.wmax = (int32_t)r.u32();     // max_quantized_value, center_value
                if (!read_rabs(r, a.aux_bits)) return UVOL_ERR_TRUNCATED;
                if (a.wmin < 3 || (a.wmin & (a.wmin + 1)) != 0) return UVOL_ERR_CORRUPT;
            } else return UVOL_ERR_UNSUPPORTED;
            if (r.err) return UVOL_ERR_TRUNCATED;
            if (a.type == 0 && f.pos_attr < 0 && a.nc == 3 && a.table < 0) f.pos_attr = d.first + j;
            // DRACOLoader exports POSITION/NORMAL/COLOR/TEX_COORD by semantic, first match wins
            // (src/lib/DRACOLoader.js:505-531); GENERIC is decoded past but not exported.
            if (a.type >= 0 && a.type <= 3 && !sem_done[a.type]) {
                sem_done[a.type] = true;
                a.out_slot = a.type == 0 ? 0 : a.type == 1 ? 1 : a.type == 3 ? 2 : 3;
            }
        }
        for (int j = 0; j < d.natt; j++) {
            DracoAttr &a = f.attr[d.first + j];
            if (a.seq == 2) { for (int k = 0; k < a.nc; k++) a.qmin[k] = r.f32(); a.qrange = r.f32(); a.qbits = r.u8(); if (a.qbits < 1 || a.qbits > 30) return UVOL_ERR_CORRUPT; }
            else if (a.seq == 3) { a.qbits = r.u8(); if (a.qbits < 2 || a.qbits > 30) return UVOL_ERR_CORRUPT; }
        }
        if (r.err) return UVOL_ERR_TRUNCATED;
    }
    if (f.pos_attr < 0) return UVOL_ERR_UNSUPPORTED;
    return UVOL_OK;
}
