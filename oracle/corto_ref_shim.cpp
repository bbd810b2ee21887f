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
#include "encoder.h"
#include "decoder.h"

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

}
