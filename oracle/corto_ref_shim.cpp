/*
 * corto_ref_shim.cpp -- thin extern "C" helpers linked INTO oracle/_ref/libcorto_ref.so next to the
 * reference's own, unmodified Corto sources (deprecated/encoder/dev/src/*.cpp, compiled in place
 * by oracle/Makefile).  TEST INFRASTRUCTURE ONLY.
 *
 * The reference exports only the decoder through its C ABI (corto_codec.h:41-43).  Tests also need
 * .crt inputs, and the repository ships none, so this shim drives the reference's public C++
 * encoder API (encoder.h:50-81: Encoder(nvert,nface,entropy), addGroup, addPositionsBits, addUvs,
 * encode, stream.data()/size()) the same way the reference CLI does (main.cpp:180-200).
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#define private public          /* the stage-wise helpers below read crt::Decoder's stream and vertex counter (test infrastructure only) */
#include "encoder.h"
#include "decoder.h"
#undef private

extern "C" {

/* Encode a triangle mesh to a malloc'd .crt blob.  uv may be NULL.  Returns size, 0 on failure.
 * out_nvert/out_nface receive the counts the encoder kept (it drops degenerate faces and
 * unreferenced vertices, encoder.cpp:247-300). */
size_t corto_ref_encode(const float *pos, const float *uv, uint32_t nvert, const uint32_t *index, uint32_t nface,
                        int pos_bits, int uv_bits, uint8_t **out, uint32_t *out_nvert, uint32_t *out_nface) {
    try {
        crt::Encoder enc(nvert, nface, crt::Stream::TUNSTALL);
        enc.addGroup((int)nface);
        std::vector<uint32_t> idx(index, index + (size_t)nface * 3);
        enc.addPositionsBits(pos, idx.data(), pos_bits);
        if (uv) enc.addUvs(uv, powf(2.0f, -(float)uv_bits));
        enc.encode();
        size_t n = enc.stream.size();
        *out = (uint8_t *)malloc(n + 4);
        memcpy(*out, enc.stream.data(), n);
        if (out_nvert) *out_nvert = enc.nvert;
        if (out_nface) *out_nface = enc.nface;
        return n;
    } catch (...) { return 0; }
}

void corto_ref_free(uint8_t *p) { free(p); }

/* Reference decode that also exposes the parallelogram context table (index.prediction), used by
 * the stage-wise parity tests.  data must be 4-byte aligned (decoder.cpp:42-43). */
int corto_ref_decode(const uint8_t *data, int len, float *pos, float *uv, uint32_t *index, uint32_t *prediction /*[nvert*3] or NULL*/) {
    try {
        crt::Decoder dec(len, data);
        if (pos) dec.setPositions(pos);
        if (uv && dec.hasAttr("uv")) dec.setUvs(uv);
        dec.setIndex(index);
        dec.decode();
        if (prediction)
            for (uint32_t i = 0; i < dec.nvert; i++) { prediction[i * 3] = dec.index.prediction[i].a; prediction[i * 3 + 1] = dec.index.prediction[i].b; prediction[i * 3 + 2] = dec.index.prediction[i].c; }
        return (int)dec.nface;
    } catch (...) { return -1; }
}


/* Full-featured encode through the reference's public API: optional uv, normals (float xyz; prediction 0 DIFF, 1 ESTIMATED,
 * 2 BORDER -- normal_attribute.h:42-44) and RGBA colours (encoder.h:59-64). */
size_t corto_ref_encode2(const float *pos, const float *uv, const float *normal, const uint8_t *color, uint32_t nvert, const uint32_t *index, uint32_t nface,
                         int pos_bits, int uv_bits, int normal_bits, int normal_pred, int color_bits,
                         uint8_t **out, uint32_t *out_nvert, uint32_t *out_nface) {
    try {
        crt::Encoder enc(nvert, nface, crt::Stream::TUNSTALL);
        enc.addGroup((int)nface);
        std::vector<uint32_t> idx(index, index + (size_t)nface * 3);
        enc.addPositionsBits(pos, idx.data(), pos_bits);
        if (uv) enc.addUvs(uv, powf(2.0f, -(float)uv_bits));
        if (normal) enc.addNormals(normal, normal_bits, (crt::NormalAttr::Prediction)normal_pred);
        if (color) enc.addColors(color, color_bits, color_bits, color_bits, color_bits);
        enc.encode();
        size_t n = enc.stream.size();
        *out = (uint8_t *)malloc(n + 4);
        memcpy(*out, enc.stream.data(), n);
        if (out_nvert) *out_nvert = enc.nvert;
        if (out_nface) *out_nface = enc.nface;
        return n;
    } catch (...) { return 0; }
}

/* Reference decode with every attribute the file carries: normals as float xyz (Decoder::setNormals(float *)), colours as RGBA8
 * (Decoder::setColors(uchar *, 4) -- the UINT8 path of ColorAttr::dequantize, color_attribute.cpp:69-90). */
int corto_ref_decode2(const uint8_t *data, int len, float *pos, float *uv, float *normal, uint8_t *color, uint32_t *index) {
    try {
        crt::Decoder dec(len, data);
        if (pos) dec.setPositions(pos);
        if (uv && dec.hasAttr("uv")) dec.setUvs(uv);
        if (normal && dec.hasAttr("normal")) dec.setNormals(normal);
        if (color && dec.hasAttr("color")) dec.setColors(color, 4);
        dec.setIndex(index);
        dec.decode();
        return (int)dec.nface;
    } catch (...) { return -1; }
}

/* Stage-wise helper for the connectivity walk: the decoded CLERS symbols and the raw index bit stream of a file, as the
 * reference's own IndexAttribute::decode (index_attribute.h:83-87) leaves them.  Returns the number of symbols (<0 on failure);
 * clers / words may be NULL to query the sizes. */
int corto_ref_clers(const uint8_t *data, int len, uint8_t *clers, int clers_cap, uint32_t *words, int words_cap, uint32_t *out_nwords, uint32_t *out_group_ends, int groups_cap, uint32_t *out_ngroups) {
    try {
        crt::Decoder dec(len, data);
        dec.index.decodeGroups(dec.stream);
        dec.index.decode(dec.stream);
        const int n = (int)dec.index.clers.size();
        if (clers && n <= clers_cap) memcpy(clers, dec.index.clers.data(), (size_t)n);
        const uint32_t nw = (uint32_t)dec.index.bitstream.size;
        if (out_nwords) *out_nwords = nw;
        if (words && (int)nw <= words_cap) memcpy(words, dec.index.bitstream.buffer, (size_t)nw * 4);
        if (out_ngroups) *out_ngroups = (uint32_t)dec.index.groups.size();
        if (out_group_ends) for (size_t g = 0; g < dec.index.groups.size() && (int)g < groups_cap; g++) out_group_ends[g] = dec.index.groups[g].end;
        return n;
    } catch (...) { return -1; }
}

}
