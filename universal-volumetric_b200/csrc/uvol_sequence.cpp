// uvol_sequence.cpp -- the C++ host side above the batch entry points: a clip on local storage opened by its MANIFEST, decoded by
// frame / segment ranges.  Mirrors the decode side of the reference's players:
//   * manifest schema, both V2 dialects and V1      src/Interfaces.ts:1-15,75-132; scripts/Encoder.py:311-328; Encoder30.js:155-160
//   * V1 / V2 dispatch on version == "v2"           src/Player.ts:127-132
//   * target choice, path templates                 src/V2/player.ts:141-174,199-221; src/utils.ts:10-45
//   * decodeDraco(url, frameNo) / decodeKTX2(url, segmentNo) keyed by frame / segment number    src/V2/player.ts:325-366
//   * V1: .manifest -> .drcs, one range read, per-frame slices                                    src/V1/player.ts:337; src/V1/worker.ts:37-56
// The decode itself is uvol_decode_v2_batch / uvol_decode_corto_batch (CUDA); nothing here decodes a payload byte.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <memory>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>
#include "uvol_ctx.h"

namespace {

// ---- a small JSON reader (objects keep their key order: the player picks the FIRST geometry target, src/V2/player.ts:207)
struct JVal {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    double num = 0; bool b = false; std::string str;
    std::vector<JVal> arr; std::vector<std::pair<std::string, JVal>> obj;
    const JVal *get(const char *k) const { if (kind != OBJ) return nullptr; for (auto &kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
    double number(const char *k, double dflt) const { const JVal *v = get(k); return v && v->kind == NUM ? v->num : dflt; }
    std::string string(const char *k, const char *dflt) const { const JVal *v = get(k); return v && v->kind == STR ? v->str : std::string(dflt); }
};
struct JParser {
    const char *p, *end; bool err = false; int depth = 0;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool lit(const char *s) { const size_t n = strlen(s); if ((size_t)(end - p) >= n && !memcmp(p, s, n)) { p += n; return true; } return false; }
    std::string parse_string() {
        std::string out;
        if (p >= end || *p != '"') { err = true; return out; }
        for (p++; p < end && *p != '"'; p++) {
            if (*p != '\\') { out.push_back(*p); continue; }
            if (++p >= end) break;
            switch (*p) {
            case 'n': out.push_back('\n'); break; case 't': out.push_back('\t'); break; case 'r': out.push_back('\r'); break;
            case 'b': out.push_back('\b'); break; case 'f': out.push_back('\f'); break;
            case 'u': {          // BMP code point -> UTF-8 (paths in manifests are ASCII in practice)
                if (end - p < 5) { err = true; return out; }
                unsigned cp = (unsigned)strtoul(std::string(p + 1, 4).c_str(), nullptr, 16); p += 4;
                if (cp < 0x80) out.push_back((char)cp);
                else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 63))); }
                else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 63))); out.push_back((char)(0x80 | (cp & 63))); }
                break; }
            default: out.push_back(*p);
            }
        }
        if (p >= end) { err = true; return out; }
        p++;
        return out;
    }
    JVal parse() {
        JVal v; ws();
        if (p >= end || ++depth > 64) { err = true; return v; }
        if (*p == '{') {
            v.kind = JVal::OBJ; p++; ws();
            if (p < end && *p == '}') { p++; depth--; return v; }
            while (!err) {
                ws(); std::string k = parse_string(); ws();
                if (err || p >= end || *p != ':') { err = true; break; }
                p++; JVal c = parse(); v.obj.emplace_back(std::move(k), std::move(c)); ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                err = true;
            }
        } else if (*p == '[') {
            v.kind = JVal::ARR; p++; ws();
            if (p < end && *p == ']') { p++; depth--; return v; }
            while (!err) {
                v.arr.push_back(parse()); ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                err = true;
            }
        } else if (*p == '"') { v.kind = JVal::STR; v.str = parse_string(); }
        else if (lit("true")) { v.kind = JVal::BOOL; v.b = true; }
        else if (lit("false")) { v.kind = JVal::BOOL; }
        else if (lit("null")) { }
        else { char *e = nullptr; v.num = strtod(p, &e); if (e == p || e > end) err = true; else { v.kind = JVal::NUM; p = e; } }
        depth--;
        return v;
    }
};

bool read_file(const std::string &path, std::vector<uint8_t> &out, uint64_t off = 0, uint64_t len = ~0ull) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    if (len == ~0ull) { fseek(f, 0, SEEK_END); const long n = ftell(f); if (n < 0) { fclose(f); return false; } len = (uint64_t)n > off ? (uint64_t)n - off : 0; }
    out.resize(len);
    bool ok = fseek(f, (long)off, SEEK_SET) == 0 && (len == 0 || fread(out.data(), 1, len, f) == len);
    fclose(f);
    return ok;
}

// src/utils.ts:10-45
std::string pad(long n, size_t width) { std::string s = std::to_string(n); return s.size() >= width ? s : std::string(width - s.size(), '0') + s; }
size_t count_hash(const std::string &s) { size_t n = 0; for (char c : s) n += c == '#'; return n; }
std::string absolute_url(const std::string &manifest, const std::string &seg) {
    if (seg.compare(0, 4, "http") == 0) return seg;
    const size_t slash = manifest.find_last_of('/');
    return slash == std::string::npos ? seg : manifest.substr(0, slash + 1) + seg;
}
void replace_first(std::string &s, const std::string &from, const std::string &to) { const size_t at = s.find(from); if (at != std::string::npos) s.replace(at, from.size(), to); }
const char *format_ext(const std::string &fmt) { return fmt == "draco" ? ".drc" : fmt == "ktx2" ? ".ktx2" : fmt == "mp3" ? ".mp3" : fmt == "etc2" ? ".etc2" : ""; }   // Interfaces.ts:156-161

struct V1Frame { uint32_t frame, keyframe; uint64_t start, length; uint32_t vertices, faces; };

}  // namespace

struct uvol_sequence {
    uvol_ctx *ctx = nullptr; std::string manifest_path; int version = 2;
    // V2
    std::string geo_target, geo_format, geo_path, tex_target, tex_format, tex_path, tex_type = "baseColor", tex_tag = "default";   // src/V2/player.ts:82-83
    double geo_fps = 0, tex_fps = 0; uint32_t geo_frames = 0, seq_size = 0, seq_count = 0;
    // V1
    std::string drcs_path; double v1_fps = 0; uint32_t max_vertices = 0, max_triangles = 0; std::vector<V1Frame> v1;
    // payload bytes of the range being decoded (kept until the next call)
    std::vector<std::vector<uint8_t>> files_a, files_b;
};

static std::string geometry_url(const uvol_sequence &s, long frame) {          // src/V2/player.ts:141-156
    std::string t = s.geo_path; const size_t w = count_hash(t);
    replace_first(t, "[target]", s.geo_target); replace_first(t, "[ext]", format_ext(s.geo_format)); replace_first(t, "[" + std::string(w, '#') + "]", pad(frame, w));
    return absolute_url(s.manifest_path, t);
}
static std::string texture_url(const uvol_sequence &s, long segment) {         // src/V2/player.ts:158-174
    std::string t = s.tex_path; const size_t w = count_hash(t);
    replace_first(t, "[target]", s.tex_target); replace_first(t, "[type]", s.tex_type); replace_first(t, "[tag]", s.tex_tag);
    replace_first(t, "[ext]", format_ext(s.tex_format)); replace_first(t, "[" + std::string(w, '#') + "]", pad(segment, w));
    return absolute_url(s.manifest_path, t);
}

// ctx may be NULL: the sequence can then be inspected (info, URLs, frame mapping) but not decoded.
#define SEQ_ERROR(ctx, msg) do { if (ctx) (ctx)->set_error(msg); } while (0)
extern "C" int uvol_open(uvol_ctx *ctx, const char *manifest_path, uvol_sequence **out) {
    if (!manifest_path || !out) return UVOL_ERR_ARG;
    *out = nullptr;
    std::vector<uint8_t> text;
    if (!read_file(manifest_path, text)) { SEQ_ERROR(ctx, "cannot read the manifest"); return UVOL_ERR_IO; }
    JParser jp{(const char *)text.data(), (const char *)text.data() + text.size()};
    const JVal m = jp.parse();
    if (jp.err || m.kind != JVal::OBJ) { SEQ_ERROR(ctx, "manifest is not a JSON object"); return UVOL_ERR_CORRUPT; }
    std::unique_ptr<uvol_sequence> s(new uvol_sequence());
    s->ctx = ctx; s->manifest_path = manifest_path;
    if (m.string("version", "") == "v2") {
        s->version = 2;
        const JVal *g = m.get("geometry"), *t = m.get("texture");
        if (!g || !t || g->kind != JVal::OBJ || t->kind != JVal::OBJ) { SEQ_ERROR(ctx, "v2 manifest without geometry / texture"); return UVOL_ERR_CORRUPT; }
        const JVal *gt = g->get("targets");
        if (gt && gt->kind == JVal::OBJ && !gt->obj.empty()) {          // the player's schema: the first target (src/V2/player.ts:207)
            s->geo_target = gt->obj[0].first; const JVal &e = gt->obj[0].second;
            s->geo_format = e.string("format", "draco"); s->geo_fps = e.number("frameRate", 0); s->geo_frames = (uint32_t)e.number("frameCount", 0);
        } else {                                                        // the encoder script's dialect (scripts/Encoder.py:311-318)
            s->geo_format = g->string("format", "draco"); s->geo_target = s->geo_format; s->geo_fps = g->number("frameRate", 0); s->geo_frames = (uint32_t)g->number("frameCount", 0);
        }
        s->geo_path = g->string("path", "");
        const JVal *tt = t->get("targets"); const JVal *pick = nullptr; std::string tpath = t->string("path", "");
        if (tt && tt->kind == JVal::OBJ && !tt->obj.empty()) {          // first key, unless a key is literally "ktx2" (playTrack's format test compares keys, :209-221)
            pick = &tt->obj[0].second; s->tex_target = tt->obj[0].first;
            for (auto &kv : tt->obj) if (kv.first == "ktx2") { pick = &kv.second; s->tex_target = kv.first; break; }
        } else if (tt && tt->kind == JVal::ARR && !tt->arr.empty()) {   // encoder dialect: a list, each entry with its own path
            pick = &tt->arr[0]; s->tex_target = pick->string("format", "ktx2");
            const std::string p = pick->string("path", ""); if (!p.empty()) tpath = p;
        }
        if (!pick || pick->kind != JVal::OBJ) { SEQ_ERROR(ctx, "v2 manifest without a texture target"); return UVOL_ERR_CORRUPT; }
        s->tex_format = pick->string("format", "ktx2"); s->tex_type = pick->string("type", "baseColor"); s->tex_tag = pick->string("tag", "default");
        s->tex_fps = pick->number("frameRate", 0); s->seq_size = (uint32_t)pick->number("sequenceSize", 0); s->seq_count = (uint32_t)pick->number("sequenceCount", 0);
        s->tex_path = tpath;
        if (s->geo_path.empty() || s->tex_path.empty() || !s->geo_frames || !s->seq_size) { SEQ_ERROR(ctx, "v2 manifest: missing path / frameCount / sequenceSize"); return UVOL_ERR_CORRUPT; }
        if (s->geo_format != "draco" || s->tex_format != "ktx2") { SEQ_ERROR(ctx, "only the draco geometry target and the ktx2 texture target decode here"); return UVOL_ERR_UNSUPPORTED; }
    } else {
        s->version = 1;                                                 // V1Schema (src/Interfaces.ts:1-15)
        const JVal *fd = m.get("frameData");
        if (!fd || fd->kind != JVal::ARR) { SEQ_ERROR(ctx, "manifest is neither v2 nor a V1 manifest with frameData"); return UVOL_ERR_CORRUPT; }
        s->v1_fps = m.number("frameRate", 0); s->max_vertices = (uint32_t)m.number("maxVertices", 0); s->max_triangles = (uint32_t)m.number("maxTriangles", 0);
        for (const JVal &e : fd->arr) {
            V1Frame f; f.frame = (uint32_t)e.number("frameNumber", 0); f.keyframe = (uint32_t)e.number("keyframeNumber", f.frame);
            f.start = (uint64_t)e.number("startBytePosition", 0); f.length = (uint64_t)e.number("meshLength", 0);
            f.vertices = (uint32_t)e.number("vertices", 0); f.faces = (uint32_t)e.number("faces", 0);
            s->v1.push_back(f);
        }
        s->drcs_path = manifest_path;                                   // `.manifest` -> `.drcs` (src/V1/player.ts:337)
        const size_t dot = s->drcs_path.rfind(".manifest");
        if (dot != std::string::npos) s->drcs_path.replace(dot, 9, ".drcs"); else s->drcs_path += ".drcs";
    }
    *out = s.release();
    return UVOL_OK;
}

extern "C" void uvol_close(uvol_sequence *seq) { delete seq; }

extern "C" int uvol_sequence_get_info(const uvol_sequence *s, uvol_sequence_info *out) {
    if (!s || !out) return UVOL_ERR_ARG;
    memset(out, 0, sizeof *out);
    out->version = s->version;
    if (s->version == 2) { out->geometry_frame_count = s->geo_frames; out->geometry_frame_rate = s->geo_fps; out->texture_frame_rate = s->tex_fps; out->sequence_size = s->seq_size; out->sequence_count = s->seq_count; }
    else { out->geometry_frame_count = (uint32_t)s->v1.size(); out->geometry_frame_rate = s->v1_fps; out->max_vertices = s->max_vertices; out->max_triangles = s->max_triangles; }
    return UVOL_OK;
}

// The file a frame / segment number maps to (for callers that fetch the bytes themselves).  Returns the length, or a negative status.
extern "C" int uvol_sequence_url(const uvol_sequence *s, int kind /*0 geometry frame, 1 texture segment*/, int number, char *buf, size_t cap) {
    if (!s || !buf || !cap || s->version != 2 || number < 0) return UVOL_ERR_ARG;
    const std::string u = kind == 0 ? geometry_url(*s, number) : texture_url(*s, number);
    if (u.size() + 1 > cap) return UVOL_ERR_ARG;
    memcpy(buf, u.c_str(), u.size() + 1);
    return (int)u.size();
}

// time -> frame / segment / layer (src/V2/player.ts:43-45,418-420,446): Math.round(fps * t), floor(texFrame / sequenceSize), texFrame % sequenceSize
extern "C" int uvol_sequence_frames_at(const uvol_sequence *s, double t, uint32_t *geometry_frame, uint32_t *segment, uint32_t *layer) {
    if (!s || s->version != 2 || !s->seq_size) return UVOL_ERR_ARG;
    const long gf = (long)floor(s->geo_fps * t + 0.5), tf = (long)floor(s->tex_fps * t + 0.5);
    if (geometry_frame) *geometry_frame = (uint32_t)gf;
    if (segment) *segment = (uint32_t)(tf / s->seq_size);
    if (layer) *layer = (uint32_t)(tf % s->seq_size);
    return UVOL_OK;
}

// getTranscoderFormat (src/lib/KTX2Loader.js:659-689) over FORMAT_OPTIONS in the order the reference effectively applies (UASTC priorities).
extern "C" int uvol_pick_texture_format(int source_is_uastc, int has_alpha, uint32_t caps) {
    if ((caps & UVOL_CAP_ASTC) && source_is_uastc) return UVOL_TEX_ASTC_4x4;                       // basisFormat: [UASTC_4x4] only
    if (caps & UVOL_CAP_BPTC) return UVOL_TEX_BC7;                                                  // [BC7_M5, BC7_M5]
    if ((caps & UVOL_CAP_ETC2) && !source_is_uastc) return has_alpha ? UVOL_TEX_ETC2_RGBA : UVOL_TEX_ETC1;      // [ETC1, ETC2]; from UASTC: not produced here
    if ((caps & UVOL_CAP_ETC1) && !source_is_uastc && !has_alpha) return UVOL_TEX_ETC1;            // [ETC1]: one entry, skipped for alpha (:672)
    if ((caps & UVOL_CAP_DXT) && !source_is_uastc) return has_alpha ? UVOL_TEX_BC3 : UVOL_TEX_BC1; // [BC1, BC3]
    return UVOL_TEX_RGBA32;                                                                         // PVRTC is not produced; the reference's own last resort
}

// fetchBuffers' leaky bucket (src/V2/player.ts:272-323), as manifest.py V2Manifest.fetch_window: for every whole second i of the look-ahead the
// request grows to min(current + (i + 1) * per-second, last index); what is new since the last request is one contiguous range.
extern "C" int uvol_sequence_fetch_window(const uvol_sequence *s, double t, int32_t *last_geometry, int32_t *last_segment, double buffer_duration_s, uvol_fetch_plan *plan) {
    if (!s || s->version != 2 || !s->seq_size || !last_geometry || !last_segment || !plan || buffer_duration_s < 0) return UVOL_ERR_ARG;
    const long gsize = (long)s->geo_fps, cur_g = (long)floor(s->geo_fps * t + 0.5);
    const long tsize = (long)ceil(s->tex_fps / (double)s->seq_size), cur_s = (long)floor(s->tex_fps * t + 0.5) / (long)s->seq_size;
    const long g_last = (long)s->geo_frames - 1, s_last = (long)s->seq_count - 1;
    long lg = *last_geometry, ls = *last_segment;
    plan->first_frame = (int32_t)(lg + 1); plan->first_segment = (int32_t)(ls + 1);
    for (long i = 0; i < (long)buffer_duration_s; i++) {
        const long g_end = std::min(cur_g + (i + 1) * gsize, g_last), s_end = std::min(cur_s + (i + 1) * tsize, s_last);
        if (lg != g_last && lg < g_end) lg = g_end;
        if (ls != s_last && ls < s_end) ls = s_end;
    }
    plan->n_frames = (int32_t)(lg - *last_geometry); plan->n_segments = (int32_t)(ls - *last_segment);
    *last_geometry = (int32_t)lg; *last_segment = (int32_t)ls;
    return UVOL_OK;
}
extern "C" int uvol_sequence_keep_from(const uvol_sequence *s, double t, int32_t *first_frame_to_keep, int32_t *first_segment_to_keep) {
    if (!s || s->version != 2 || !s->seq_size || s->geo_fps <= 0 || s->tex_fps <= 0) return UVOL_ERR_ARG;
    const long gf = (long)floor(s->geo_fps * t + 0.5), seg = (long)floor(s->tex_fps * t + 0.5) / (long)s->seq_size;
    if (first_frame_to_keep) *first_frame_to_keep = (int32_t)(gf - (long)ceil(120.0 / s->geo_fps));                                  // screens up to 120 Hz (:544-546)
    if (first_segment_to_keep) *first_segment_to_keep = (int32_t)(seg - (long)ceil(120.0 / (s->tex_fps * (double)s->seq_size)));
    return UVOL_OK;
}

// V2: geometry frames [first_frame, +n_frames) and texture segments [first_segment, +n_segments) read from the files the manifest names
// and decoded in one uvol_decode_v2_batch call (either count may be 0).  A file that cannot be read is a per-item UVOL_STATUS_IO.
extern "C" int uvol_decode_range(uvol_sequence *s, int first_frame, int n_frames, int first_segment, int n_segments, int memory, uvol_geometry *out_geo, uvol_texture *out_tex) {
    if (!s || !s->ctx || s->version != 2 || first_frame < 0 || n_frames < 0 || first_segment < 0 || n_segments < 0 || (n_frames && !out_geo) || (n_segments && !out_tex)) return UVOL_ERR_ARG;
    if ((uint64_t)first_frame + n_frames > s->geo_frames || (s->seq_count && (uint64_t)first_segment + n_segments > s->seq_count)) return UVOL_ERR_ARG;
    s->files_a.assign((size_t)n_frames, {}); s->files_b.assign((size_t)n_segments, {});
    std::vector<const uint8_t *> pa((size_t)n_frames), pb((size_t)n_segments); std::vector<size_t> sa((size_t)n_frames), sb((size_t)n_segments);
    std::vector<char> bad_a((size_t)n_frames, 0), bad_b((size_t)n_segments, 0);
    for (int i = 0; i < n_frames; i++) { bad_a[i] = !read_file(geometry_url(*s, first_frame + i), s->files_a[i]); pa[i] = bad_a[i] ? nullptr : s->files_a[i].data(); sa[i] = s->files_a[i].size(); }
    for (int i = 0; i < n_segments; i++) { bad_b[i] = !read_file(texture_url(*s, first_segment + i), s->files_b[i]); pb[i] = bad_b[i] ? nullptr : s->files_b[i].data(); sb[i] = s->files_b[i].size(); }
    const int rc = uvol_decode_v2_batch(s->ctx, pa.data(), sa.data(), n_frames, pb.data(), sb.data(), n_segments, memory, out_geo, out_tex);
    if (rc) return rc;
    for (int i = 0; i < n_frames; i++) if (bad_a[i]) out_geo[i].status = UVOL_ERR_IO;
    for (int i = 0; i < n_segments; i++) if (bad_b[i]) out_tex[i].status = UVOL_ERR_IO;
    return UVOL_OK;
}

// V1: frames [first, first + n) of the manifest's frameData -- one range read of the .drcs, per-frame slices, one batched decode
// (src/V1/worker.ts:37-68).  out[i] belongs to frameData[first + i]; keyframe_numbers (optional) receives the keys the player's
// meshBuffer uses (src/V1/player.ts:296).
extern "C" int uvol_decode_v1_range(uvol_sequence *s, int first, int n, int memory, uvol_corto_mesh *out, uint32_t *keyframe_numbers) {
    if (!s || !s->ctx || s->version != 1 || first < 0 || n < 0 || (n && !out) || (uint64_t)first + n > s->v1.size()) return UVOL_ERR_ARG;
    if (n == 0) return UVOL_OK;
    const uint64_t lo = s->v1[first].start, hi = s->v1[first + n - 1].start + s->v1[first + n - 1].length;
    s->files_a.assign(1, {});
    if (hi < lo || !read_file(s->drcs_path, s->files_a[0], lo, hi - lo)) { s->ctx->set_error("cannot read the .drcs range"); return UVOL_ERR_IO; }
    std::vector<const uint8_t *> p((size_t)n); std::vector<size_t> sz((size_t)n);
    for (int i = 0; i < n; i++) {
        const V1Frame &f = s->v1[first + i];
        const bool ok = f.start >= lo && f.start + f.length <= hi;
        p[i] = ok ? s->files_a[0].data() + (f.start - lo) : nullptr; sz[i] = ok ? (size_t)f.length : 0;
        if (keyframe_numbers) keyframe_numbers[i] = f.keyframe;
    }
    return uvol_decode_corto_batch(s->ctx, p.data(), sz.data(), n, memory, out);
}
