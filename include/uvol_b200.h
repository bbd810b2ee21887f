/*
 * uvol_b200.h -- C ABI of libuvol_b200.so, the B200-native UVOL decode path.
 *
 * The reference has no single FFI for its V2 path (JS -> WASM inside Web Workers); its contract is
 * the manifest plus the worker message payloads.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference repo).  Plain pointers and sizes only.
 *
 * Ownership: all output arrays are LIBRARY-owned (mirrors the transfer-of-ownership the workers
 * do with transferable ArrayBuffers, src/lib/DRACOLoader.js:152,449; src/lib/KTX2Loader.js:335,431).
 * Buffers returned by a *_batch call stay valid until the next *_batch call of the same kind on
 * the same ctx, or uvol_destroy.  With UVOL_MEM_DEVICE they are device pointers on the ctx's GPU,
 * with UVOL_MEM_HOST pinned host pointers (the device->host copy is part of the call).
 *
 * Errors: functions return 0 or a negative uvol_status; per-item failures are reported in
 * item.status and never abort the batch (mirrors "a failed frame is simply absent from meshMap",
 * src/V2/player.ts:429-444, and the worker's {type:'error'} reply, DRACOLoader.js:451-455).
 * Threading: one ctx per GPU; calls on one ctx must be serialised by the caller; distinct ctxs are
 * independent (mirrors one decoder instance per worker, DRACOLoader.js:439).
 * There is no CPU decode path: without a usable CUDA device every call fails with UVOL_ERR_CUDA.
 */
#ifndef UVOL_B200_H
#define UVOL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uvol_ctx uvol_ctx;

enum uvol_status { UVOL_STATUS_OK = 0, UVOL_STATUS_TRUNCATED = -1, UVOL_STATUS_CORRUPT = -2, UVOL_STATUS_UNSUPPORTED = -3,
                   UVOL_STATUS_CUDA = -4, UVOL_STATUS_ARG = -5, UVOL_STATUS_IO = -6 };
enum uvol_memory { UVOL_MEM_DEVICE = 0, UVOL_MEM_HOST = 1 };
/* Target texture format (the reference picks it from the GPU's capabilities, getTranscoderFormat / FORMAT_OPTIONS,
 * src/lib/KTX2Loader.js:591-689).  RGBA32 is the parity target and the reference's own fallback (:682-687).  ETC1 is the
 * `etc1Supported` / opaque `etc2Supported` choice (:619-636; an RGB ETC2 texture of ETC1S content is its ETC1 blocks): 8 bytes per
 * 4x4 block in block raster order, layers back to back; opaque ETC1S sources only (others report UVOL_STATUS_UNSUPPORTED per item).
 * BC7 is what those options select for ETC1S *and* UASTC sources on a desktop GPU with EXT_texture_compression_bptc (BC7_M5 /
 * RGBA_BPTC_Format, :602-604): 16 bytes per 4x4 block in block raster order, layers back to back; ETC1S (with or without alpha) ->
 * BC7 mode 5, UASTC -> the BC7 mode with the same subset shapes and weight grid (csrc/bc7_core.h).  Not bit-identical to the
 * reference's transcoder (whose tables are not in its tree): validated by an independent BC7 decoder, bounds in tests/test_bc7.py.
 * ASTC_4x4 is the reference's first choice for UASTC sources on GPUs with WEBGL_compressed_texture_astc (ASTC_4x4 /
 * RGBA_ASTC_4x4_Format, :592-600; not offered for ETC1S: those items report UVOL_STATUS_UNSUPPORTED): 16 bytes per block, block raster
 * order, layers back to back.  UASTC is a subset of ASTC, so this repack is LOSSLESS: the blocks decode to exactly the RGBA32 texels
 * (csrc/astc_core.h; checked bit for bit through an independent ASTC decoder, tests/test_astc.py). */
enum uvol_texture_format { UVOL_TEX_RGBA32 = 0, UVOL_TEX_ETC1 = 1, UVOL_TEX_BC7 = 2,
                           UVOL_TEX_ETC2_RGB = 3 /* only as the format of uvol_upload_etc2_batch results: raw blocks passed through */,
                           UVOL_TEX_ASTC_4x4 = 4,
                           UVOL_TEX_ETC2_RGBA = 5 /* `etc2Supported` with alpha (ETC2 / RGBA_ETC2_EAC_Format, :619-627, the top-priority option for ETC1S
                             sources): 16 bytes per block = EAC alpha block + the ETC1 colour block (exact); ETC1S sources only, opaque ones get a
                             constant-255 alpha block; the alpha fit is lossy (<= 12 / 255 per texel, >= 42 dB; csrc/basis_core.h, tests/test_etc2.py) */,
                           UVOL_TEX_BC1 = 6, UVOL_TEX_BC3 = 7 /* `dxtSupported` (BC1 / BC3, RGB_S3TC_DXT1 / RGBA_S3TC_DXT5, :610-618), the fallback on desktop GPUs
                             without BPTC: 8 bytes per block (BC1) or the BC4 alpha block + the BC1 block (BC3); ETC1S sources only; lossy (RGB565
                             endpoints), decoded by Pillow's DXT decoder in tests/test_dxt.py */ };

/* getTranscoderFormat (src/lib/KTX2Loader.js:659-689): the target a context with the given capabilities gets for a source.  The options are
 * tried in the order the reference effectively uses for BOTH source kinds (its two option lists alias one array sorted in place twice,
 * :648-657, so the UASTC priorities win): ASTC (UASTC sources only) -> BC7 -> ETC2 pair [ETC1, ETC2 RGBA] -> ETC1 (opaque only) -> DXT pair
 * [BC1, BC3] -> PVRTC, else RGBA32 (:682-687).  An option this library cannot produce for that source (PVRTC; ETC / DXT from UASTC) is passed
 * over like an unsupported capability, so the result is always a format uvol_transcode_ktx2_batch accepts for the file. */
/* The KTX2File getters the reference reads before it picks a target (getWidth / getHeight / getLayers / getLevels / getFaces / getHasAlpha /
 * isUASTC / isVideo, src/lib/KTX2Loader.js:471-495): container header only, host code, nothing is decoded.  Returns 0 or the status the
 * transcode call would report for the file. */
typedef struct uvol_ktx2_info { uint32_t width, height, layers, levels, faces, is_uastc, has_alpha, is_video, supercompression, dfd_transfer, dfd_flags; } uvol_ktx2_info;
int uvol_ktx2_probe(const uint8_t *data, size_t size, uvol_ktx2_info *out);
enum uvol_gpu_caps { UVOL_CAP_ASTC = 1, UVOL_CAP_BPTC = 2, UVOL_CAP_DXT = 4, UVOL_CAP_ETC2 = 8, UVOL_CAP_ETC1 = 16, UVOL_CAP_PVRTC = 32 };
int uvol_pick_texture_format(int source_is_uastc, int has_alpha, uint32_t caps);

/* Result of one geometry frame.  Replaces the Draco worker reply
 *   {type:'decode', geometry:{index:{array:Uint32Array(F*3)}, attributes:[{name, array:Float32Array(P*itemSize), itemSize}]}}
 * (src/lib/DRACOLoader.js:449,502,567,584-588); attribute presence follows the semantic lookup at
 * :523-525 (POSITION, NORMAL, COLOR, TEX_COORD; GENERIC ignored).  All arrays are in Draco point order. */
typedef struct uvol_geometry {
    int32_t status;
    uint32_t num_points;
    uint32_t num_faces;
    uint32_t color_components;
    uint32_t *index;       /* u32[num_faces*3] */
    float *position;       /* f32[num_points*3] */
    float *normal;         /* f32[num_points*3] or NULL */
    float *uv;             /* f32[num_points*2] or NULL */
    float *color;          /* f32[num_points*color_components] or NULL */
} uvol_geometry;

/* Result of one KTX2 segment.  Replaces the Basis worker reply
 *   {type:'transcode', faces:[{mipmaps:[{data, width, height}], ...}], width, height, hasAlpha, format, dfdTransferFn, dfdFlags}
 * (src/lib/KTX2Loader.js:431,565-578): `data` holds all layers back to back (concat, :565). */
typedef struct uvol_texture_level {   /* one entry of `mipmaps` (KTX2Loader.js:514-573): all layers of that level back to back at data + offset */
    uint32_t width, height;            /* max(1, base >> level) */
    uint64_t offset, bytes;
} uvol_texture_level;
typedef struct uvol_texture {
    int32_t status;
    uint32_t width, height, layers;
    uint32_t format;       /* uvol_texture_format */
    uint32_t has_alpha, dfd_transfer, dfd_flags;
    uint8_t *data;         /* level 0.  RGBA32: u8[layers * width * height * 4]; ETC1: u8[layers * ceil(w/4) * ceil(h/4) * 8]; BC1: same * 8; BC7, ASTC_4x4, ETC2_RGBA, BC3: ... * 16 */
    uint64_t bytes;        /* level 0 when levels == 1; with a mip chain: up to the end of the last level (levels are 128-byte aligned) */
    uint32_t levels;       /* levelCount of the file (1 for UVOL content: scripts/Encoder.py writes no mips); cube faces are not supported */
    uint32_t reserved;
    const uvol_texture_level *mips;   /* [levels], library-owned like data; mips[0].offset == 0 (NULL for uvol_upload_etc2_batch results) */
} uvol_texture;

/* Timing / traffic of the last batch call on a ctx (CUDA events on the ctx's stream). */
typedef struct uvol_stats {
    double host_parse_ms, h2d_ms, device_ms, d2h_ms, total_ms;
    float stage_ms[24];            /* per-kernel-stage device time; names via uvol_stage_name() */
    uint32_t num_stages, kernel_launches;
    uint64_t bytes_in, bytes_out;  /* compressed bytes consumed / final output bytes produced */
    uint64_t scratch_bytes;
} uvol_stats;

/* Tunables of a context (SURVEY 5 "config / flags").  The reference takes constructor arguments only (bufferDuration = 4,
 * intervalDuration = 2, src/Player.ts:50-51) and derives the texture target from the GPU's capabilities
 * (src/lib/KTX2Loader.js:591-689); here they are one struct.  uvol_config_default() fills the defaults and then applies the
 * environment overrides UVOL_TEXTURE_TARGET (rgba32 | etc1 | bc7 | astc | etc2 | bc1 | bc3), UVOL_CORTO_INDEX_U16, UVOL_STAGING_THREADS, UVOL_MAX_FACES,
 * UVOL_MAX_TEXTURE_BYTES, UVOL_BUFFER_DURATION, UVOL_INTERVAL_DURATION. */
typedef struct uvol_config {
    uint32_t struct_size;            /* sizeof(uvol_config) */
    uint32_t texture_target;         /* uvol_texture_format of uvol_decode_v2_batch and of sequences opened with uvol_open */
    uint32_t corto_index_u16;        /* V1: narrow the index to u16 when nface < 65536, the web player's layout (src/V1/player.ts:292, corto.ts:675-680) */
    uint32_t staging_threads;        /* host threads staging a large batch into pinned memory; 0 = min(8, cores / 2) */
    uint64_t max_faces_per_frame;    /* resource limit per .drc / .crt (default 2^24): larger headers fail per item with UVOL_STATUS_UNSUPPORTED */
    uint64_t max_texture_bytes;      /* resource limit per .ktx2 segment, decoded bytes (default 2^31) */
    double buffer_duration_s;        /* playback: seconds decoded ahead of the clock (default 4) */
    double interval_duration_s;      /* playback: seconds between prefetch rounds (default 2) */
} uvol_config;
void uvol_config_default(uvol_config *cfg);

/* ---- context ------------------------------------------------------------------------------- */
int uvol_create(int device, uvol_ctx **out);                                  /* = uvol_create_with_config(device, NULL, out) */
int uvol_create_with_config(int device, const uvol_config *cfg, uvol_ctx **out);
int uvol_get_config(const uvol_ctx *ctx, uvol_config *out);
void uvol_destroy(uvol_ctx *ctx);
const char *uvol_last_error(const uvol_ctx *ctx);
int uvol_get_stats(const uvol_ctx *ctx, uvol_stats *out);
const char *uvol_stage_name(int kind /*0 geometry, 1 texture, 2 corto*/, int stage);
int uvol_set_profiling(uvol_ctx *ctx, int enable);   /* per-stage CUDA events on/off (default off) */

/* ---- V2 geometry: replaces DRACOLoader.decodeGeometry -> DRACOWorker 'decode'
 * (src/lib/DRACOLoader.js:104-187,433-457,470-554) for n files at once; V2Player.decodeDraco
 * (src/V2/player.ts:325-331) maps frame numbers to files.  data[i]/size[i] = the bytes of one .drc. */
int uvol_decode_draco_batch(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n,
                            int memory, uvol_geometry *out);

/* Re-runs the device pipeline on the batch still resident in HBM from the last
 * uvol_decode_draco_batch call on this ctx (no parse, no input upload); n must match.  Measurement aid:
 * times the kernels with inputs already in HBM. */
int uvol_replay_draco_batch(uvol_ctx *ctx, int memory, uvol_geometry *out, int n);

/* ---- V2 texture: replaces KTX2Loader._createTexture -> BasisWorker.transcode
 * (src/lib/KTX2Loader.js:297-337,469-580) for n .ktx2 segments at once; V2Player.decodeKTX2
 * (src/V2/player.ts:359-366) maps segment numbers to files. */
int uvol_transcode_ktx2_batch(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n,
                              int target_format, int memory, uvol_texture *out);

int uvol_replay_ktx2_batch(uvol_ctx *ctx, int memory, uvol_texture *out, int n);

/* ---- one V2 playback step: n_drc geometry frames and n_ktx2 texture segments decoded CONCURRENTLY
 * (separate CUDA streams), like V2Player.fetchBuffers issuing decodeDraco and decodeKTX2 requests to its
 * two worker pools at once (src/V2/player.ts:272-323).  Results as for the two single-kind calls. */
int uvol_decode_v2_batch(uvol_ctx *ctx, const uint8_t *const *drc, const size_t *drc_size, int n_drc,
                         const uint8_t *const *ktx2, const size_t *ktx2_size, int n_ktx2,
                         int memory, uvol_geometry *out_geo, uvol_texture *out_tex);
int uvol_replay_v2_batch(uvol_ctx *ctx, int memory, uvol_geometry *out_geo, int n_drc, uvol_texture *out_tex, int n_ktx2);
int uvol_get_stats_kind(const uvol_ctx *ctx, int kind /*0 geometry, 1 texture*/, uvol_stats *out);

/* ---- V1 geometry: replaces the worker's per-frame `new CortoDecoder(slice).decode()` loop (src/V1/worker.ts:48-68;
 * src/lib/corto.ts:73-140 == crt::Decoder::decode, deprecated/encoder/dev/src/decoder.cpp:122-173) for n frames sliced
 * out of a .drcs by the manifest's startBytePosition / meshLength (src/Interfaces.ts:1-8).  Result = the
 * bufferGeometry of src/V1/player.ts:289-297: index (u32; with uvol_config.corto_index_u16 also narrowed to u16 when
 * nface < 65536 like the JS path, corto.ts:675-680), position f32[V*3], uv f32[V*2], plus normals / colours when the file
 * carries them.  The single-frame C ABI of the reference is in corto_codec.h. */
typedef struct uvol_corto_mesh {
    int32_t status;
    uint32_t num_vertices, num_faces;
    uint32_t index_type;   /* 0: only `index` (u32); 1: `index16` is filled as well (uvol_config.corto_index_u16 and num_faces < 65536) */
    uint32_t *index;       /* u32[num_faces*3] */
    float *position;       /* f32[num_vertices*3] */
    float *uv;             /* f32[num_vertices*2] or NULL */
    float *normal;         /* f32[num_vertices*3] or NULL: "normal" attribute, codec 2 (src/lib/corto.ts:470-671) */
    uint8_t *color;        /* u8[num_vertices*4] RGBA or NULL: "color" attribute, codec 3 (src/lib/corto.ts:439-466) */
    uint16_t *index16;     /* u16[num_faces*3] or NULL: the web player's index layout (src/V1/player.ts:292) */
} uvol_corto_mesh;
int uvol_decode_corto_batch(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int memory, uvol_corto_mesh *out);

/* ---- a clip on local storage, opened by its manifest (csrc/uvol_sequence.cpp).  Mirrors the decode side of the reference's
 * players: V1 / V2 dispatch on version == "v2" (src/Player.ts:127-132); the V2 schema in both dialects (src/Interfaces.ts:75-132;
 * scripts/Encoder.py:311-328), target choice and path templates (src/V2/player.ts:141-174,199-221; src/utils.ts:10-45);
 * decodeDraco(url, frameNo) / decodeKTX2(url, segmentNo) keyed by frame / segment number (:325-366); V1: .manifest -> .drcs, one
 * range read, per-frame slices (src/V1/player.ts:337; src/V1/worker.ts:37-68).  Results are library-owned like those of the batch
 * calls (valid until the next decode on the same ctx, or uvol_release). */
typedef struct uvol_sequence uvol_sequence;
typedef struct uvol_sequence_info {
    int32_t version;                 /* 1 or 2 */
    uint32_t geometry_frame_count;   /* V2: geometry target frameCount; V1: frameData entries */
    double geometry_frame_rate, texture_frame_rate;
    uint32_t sequence_size, sequence_count;      /* V2: layers per KTX2 segment, number of segments */
    uint32_t max_vertices, max_triangles;        /* V1 */
} uvol_sequence_info;
int uvol_open(uvol_ctx *ctx, const char *manifest_path, uvol_sequence **out);
void uvol_close(uvol_sequence *seq);
int uvol_sequence_get_info(const uvol_sequence *seq, uvol_sequence_info *out);
/* the file a geometry frame (kind 0) / texture segment (kind 1) number maps to; returns its length or a negative status */
int uvol_sequence_url(const uvol_sequence *seq, int kind, int number, char *buf, size_t cap);
/* time -> geometry frame, texture segment, layer inside the segment (src/V2/player.ts:43-45,418-420,446) */
int uvol_sequence_frames_at(const uvol_sequence *seq, double t, uint32_t *geometry_frame, uint32_t *segment, uint32_t *layer);
/* Playback planning, the decode side of V2Player.fetchBuffers' leaky bucket (src/V2/player.ts:272-323): given the clock t and the last
 * frame / segment already requested (*last_geometry / *last_segment, -1 at the start; updated), the contiguous ranges that bring the
 * buffer up to `buffer_duration_s` whole seconds ahead -- to be handed to uvol_decode_range as ONE batch instead of one worker request per
 * file.  Either count may come back 0.  uvol_sequence_keep_from: what removePlayedBuffer keeps at time t (:531-562: frames / segments
 * older than ceil(120 / fps) behind the clock may be dropped by the caller's mesh / texture maps). */
typedef struct uvol_fetch_plan { int32_t first_frame, n_frames, first_segment, n_segments; } uvol_fetch_plan;
int uvol_sequence_fetch_window(const uvol_sequence *seq, double t, int32_t *last_geometry, int32_t *last_segment, double buffer_duration_s,
                               uvol_fetch_plan *plan);
int uvol_sequence_keep_from(const uvol_sequence *seq, double t, int32_t *first_frame_to_keep, int32_t *first_segment_to_keep);
/* V2: frames [first_frame, +n_frames) and segments [first_segment, +n_segments) in one uvol_decode_v2_batch call (texture target =
 * uvol_config.texture_target); a missing file is a per-item UVOL_STATUS_IO */
int uvol_decode_range(uvol_sequence *seq, int first_frame, int n_frames, int first_segment, int n_segments, int memory,
                      uvol_geometry *out_geo, uvol_texture *out_tex);
/* V1: frameData[first .. first + n) in one uvol_decode_corto_batch call; keyframe_numbers (optional) = the keys of the player's meshBuffer */
int uvol_decode_v1_range(uvol_sequence *seq, int first, int n, int memory, uvol_corto_mesh *out, uint32_t *keyframe_numbers);
/* Frees the result buffers of `ctx` (device and pinned host) -- the explicit end of the ownership the library holds over returned
 * arrays (the reference's workers hand ownership over with transferables); every pointer returned before becomes invalid. */
int uvol_release(uvol_ctx *ctx);

/* ---- the texture-side pieces outside the KTX2 path (csrc/v1_texture.cu)
 * V1: frame number of n decoded RGBA8 video frames from the 16-cell binary counter in their bottom-left corner
 * (drawVideoAndGetCurrentFrameNumber, src/V1/player.ts:305-334; window_size = encoderWindowSize 8, byte_length = encoderByteLength 16,
 * src/Player.ts:47-48).  frames: device memory when frames_on_device != 0 (e.g. NVDEC output), else host memory. */
int uvol_v1_frame_numbers(uvol_ctx *ctx, const uint8_t *frames, int frames_on_device, int n, int width, int height,
                          int window_size, int byte_length, int32_t *out);
/* V2 'etc2' texture target (src/V2/player.ts:338-356,454-470): raw RGB-ETC2 block files, one per frame, handed to the GPU as they are.
 * Each file must hold exactly ceil(w/4) * ceil(h/4) * 8 bytes (else a per-item status); out[i].format = UVOL_TEX_ETC2_RGB. */
int uvol_upload_etc2_batch(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int width, int height,
                           int memory, uvol_texture *out);

/* Makes `ctx` use the phase-2 geometry scratch of `owner` (same device) instead of allocating its own.  For sequences
 * decoded in WINDOWS (the prefetch window of src/V2/player.ts:272-323, `fps x bufferDuration` frames): one ctx per window
 * keeps that window's compressed inputs, phase-1 scratch and outputs; the largest arena (about three quarters of a batch's
 * scratch) exists once.  Calls on the sharing contexts may be issued concurrently from one host thread per ctx: the arena is
 * handed over under a mutex, so window k+1's entropy / connectivity stages and window k's result copy overlap the other
 * window's phase 2 (mirrors the reference keeping several decode requests in flight across its worker pool). */
int uvol_share_arenas(uvol_ctx *ctx, uvol_ctx *owner);

/* Makes `ctx` return its UVOL_MEM_HOST results in the pinned host buffers of `owner` (bounds the pinned memory of a windowed
 * sequence to one window's outputs).  Results of either ctx are then valid until the next UVOL_MEM_HOST call on the other; such
 * calls must be serialised by the caller. */
int uvol_share_host_outputs(uvol_ctx *ctx, uvol_ctx *owner);

/* Device time (ms, CUDA events) spanned by the last calls of `n` contexts that ran concurrently: first kernel of any of them
 * to last kernel of any of them.  Measurement aid; needs uvol_set_profiling(ctx, 1). */
int uvol_span_ms(uvol_ctx *const *ctxs, int n, float *ms);

/* Zstandard frame decoder (RFC 8878) used for KTX2 supercompressionScheme 2 (replaces src/lib/zstddec.module.js, the WASM zstd the
 * reference inflates levels with, src/lib/KTX2Loader.js:803-817).  Host code: inflates every frame in src[0..n) into dst[0..cap);
 * *out_len receives the byte count.  No dictionaries; the content checksum is not verified. */
int uvol_zstd_inflate(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len);
/* Test aid: how often each part of the format was decoded since the last reset (see csrc/zstd_inflate.cpp for the 16 slots). */
void uvol_zstd_feature_counts(uint64_t *out16, int reset);

/* Writes a buffer larger than L2 (256 MiB) on the ctx's stream and waits: L2 flush between timed iterations. */
int uvol_flush_l2(uvol_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
