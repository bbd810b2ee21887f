/*
 * oracle.h -- C interface of the CPU oracles (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load liboracle.so.  The product library (universal-volumetric_b200/) never does.
 */
#ifndef UVOL_ORACLE_H
#define UVOL_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { UVO_OK = 0, UVO_ERR_TRUNCATED = -1, UVO_ERR_CORRUPT = -2, UVO_ERR_UNSUPPORTED = -3 };

/* ---- Draco (V2 geometry).  Output arrays follow the worker message of
 * src/lib/DRACOLoader.js:449,502,567,584-588: index u32[F*3], attributes f32[P*itemSize]. */
typedef struct {
    int status;
    uint32_t num_faces, num_points, num_vertices, num_symbols;
    uint32_t *index;          /* [3F] point ids */
    float *position;          /* [3P] or NULL */
    float *normal;            /* [3P] or NULL */
    float *uv;                /* [2P] or NULL */
    float *color;             /* [ncP] or NULL */
    /* self-consistency oracles, SURVEY.md A.4 */
    uint32_t symhist[5];      /* C S L R E */
    uint32_t attr_vertices[4];
    int ctx_counters_zero, rans_terminal_ok;
    size_t bytes_consumed;
    int pos_entries, uv_entries, nrm_entries;
    int pos_wraps, uv_wraps, uv_orient_left, nrm_flips, pos_parallelograms;
    int32_t pos_wmin, pos_wmax, uv_wmin, uv_wmax;
    /* intermediates for stage-wise parity tests */
    int32_t *pos_q, *uv_q, *nrm_q;   /* entry-order quantised values */
    int32_t *dbg_c2v, *dbg_opp;      /* [3F] base corner table */
} uvo_draco_mesh;

int uvo_draco_decode(const uint8_t *data, size_t len, uvo_draco_mesh *out);
void uvo_draco_free(uvo_draco_mesh *m);

/* ---- KTX2 / Basis (V2 texture).  Output follows src/lib/KTX2Loader.js:565-578:
 * one buffer per mip, all layers concatenated layer-major, RGBA32. */
typedef struct {
    int status;
    uint32_t width, height, layers, levels, faces;
    int is_uastc, is_video, has_alpha, dfd_transfer, dfd_flags;
    uint8_t *rgba;            /* [layers * w * h * 4], level 0 */
    size_t rgba_bytes;
    /* B.4 oracles */
    uint32_t endpoint_count, selector_count;
    uint32_t endpoints_bytes, endpoints_used, selectors_bytes, selectors_used, tables_bytes, tables_used;
    uint32_t slices, slices_exact;   /* slices whose byte length was consumed exactly */
    uint32_t pred_hist[4];
    /* intermediates */
    uint16_t *endpoint_idx, *selector_idx;   /* [layers * blocks] */
} uvo_ktx2_image;

int uvo_ktx2_decode(const uint8_t *data, size_t len, uvo_ktx2_image *out);
void uvo_ktx2_free(uvo_ktx2_image *m);

/* ---- BC7 (BPTC) decoder, all eight modes: the independent check of the product's UVOL_TEX_BC7 target (bc7_decode.c). */
int uvo_bc7_decode_block(const uint8_t *block16, uint8_t *rgba64);
int uvo_bc7_decode_image(const uint8_t *blocks, uint32_t w, uint32_t h, uint8_t *rgba);

/* ---- ASTC LDR 4x4 decoder written from the ASTC specification: the independent check of the product's UVOL_TEX_ASTC_4x4 target
 * (astc_decode.c).  decode_block: 0 = decoded, 1 = a block outside the decoder's scope; decode_image returns the number of those. */
int uvo_astc_decode_block(const uint8_t *block16, uint8_t *rgba64);
int uvo_astc_decode_image(const uint8_t *blocks, uint32_t w, uint32_t h, uint8_t *rgba);

/* ---- thread-pooled batch drivers for the CPU baseline (frames / segments are independent,
 * mirroring the <=4-worker design of DRACOLoader.js:24,312-364 and WorkerPool.js:7).
 * Return the number of items that decoded OK; checksum (optional) gets an FNV over outputs. */
int uvo_draco_decode_batch(const uint8_t *const *data, const size_t *len, int n, int threads, uint64_t *checksum,
                           uint64_t *total_points, uint64_t *total_faces);
int uvo_ktx2_decode_batch(const uint8_t *const *data, const size_t *len, int n, int threads, uint64_t *checksum,
                          uint64_t *total_texels);

#ifdef __cplusplus
}
#endif
#endif
