// uvol_internal.h -- descriptors shared by the host parsers and the sm_100a kernels.
//
// Layout rules (DESIGN.md "Data layout in HBM"):
//   * one batch = one contiguous byte blob of all input files (pinned host -> device, one copy),
//     one descriptor per frame / KTX2 file, one scratch arena addressed by 64-bit byte offsets;
//   * every per-frame array is 128-byte aligned so warps issue whole-line accesses;
//   * serial stages (entropy streams, connectivity walk, traversal, prediction chains) run one
//     warp per unit with the batch supplying the parallelism; everything else is element-parallel.
#pragma once
#include <stdint.h>
#include <stddef.h>

#define UVOL_MAX_ATTR_DATA 4      // non-position attribute connectivities per mesh
#define UVOL_MAX_ATTRS 6          // attributes per mesh (all decoders flattened)

enum { UVOL_OK = 0, UVOL_ERR_TRUNCATED = -1, UVOL_ERR_CORRUPT = -2, UVOL_ERR_UNSUPPORTED = -3, UVOL_ERR_CUDA = -4,
       UVOL_ERR_ARG = -5, UVOL_ERR_IO = -6,
       // internal (never returned to the caller): a frame outgrew its optimistic capacity / the batch outgrew the count-sized arenas.
       // The launcher re-plans (full bounds for that frame / exact arena sizes) and runs the batch again.
       UVOL_ERR_FRAME_CAPACITY = -100, UVOL_ERR_BATCH_CAPACITY = -101 };

// ---- Draco ---------------------------------------------------------------------------------
struct RansStream {          // one RAW rANS symbol run (SURVEY A.2 DecodeSymbols)
    uint32_t data_off;       // byte offset inside the frame's file
    uint32_t data_len;
    uint32_t prob_off;       // u32 offset into the batch "aux" table area: alphabet probabilities
    uint32_t alphabet;
    uint32_t count;          // symbols to decode; 0xFFFFFFFF = filled on device (entries * comps)
    uint32_t pb;             // rANS precision bits
    uint32_t nnz;            // symbols with a non-zero probability
};
struct RabsStream { uint32_t data_off, data_len, prob_zero, pad; };

struct DracoAttr {
    int8_t type, dtype, nc, normalized;       // GeometryAttribute type: 0 POSITION 1 NORMAL 2 COLOR 3 TEX_COORD 4 GENERIC
    int8_t seq;                               // 1 INTEGER 2 QUANTIZATION 3 NORMALS
    int8_t pred, xform;                       // prediction method / transform ids (A.1)
    int8_t table;                             // -1: base corner table, else attribute-data id
    int32_t vnc;                              // portable components (2 for NORMALS)
    RansStream sym;
    int32_t wmin, wmax;                       // WRAP bounds, or (max_q, center) for octahedron
    RabsStream aux_bits;                      // TEX_COORDS orientations / GEOMETRIC_NORMAL flips
    int32_t num_orient;
    uint32_t tagged, tag_bits_off, tag_bits_len;   // TAGGED symbol scheme: `sym` is the run of per-entry bit-length tags, the raw bit fields follow it in the file
    float qmin[4]; float qrange; int32_t qbits;
    int32_t out_slot;                         // 0 position 1 normal 2 uv 3 color, -1 not exported
};

struct DracoFrame {
    uint64_t file_off; uint32_t file_len; int32_t status;
    uint32_t trav, nv_enc, nf, nad, nsym, nsplit, nts, ts_off;   // ts_off: aux u32 triples {src, split, edge}
    uint32_t stdsym_off, stdsym_len;
    RabsStream start_faces, seams[UVOL_MAX_ATTR_DATA];
    RansStream ctx[6];
    int32_t nattr; DracoAttr attr[UVOL_MAX_ATTRS];
    int32_t pos_attr;                          // index of the POSITION attribute (parent of uv/normal predictors)
    uint32_t full_cap, pad0;                   // 1: size the attribute-table arrays for the full bound (one attribute vertex per corner)
    // ---- header-sized arenas S (uninitialised) and Z (zeroed per run): byte offsets filled by the host planner (draco_plan.h).
    // S = [long-lived arrays][union region: connectivity temporaries | traversal records | prediction parents]
    uint64_t o_opp, o_c2v, o_lmc, o_hole, o_val, o_stack, o_ctxsym[6], o_invalid;
    uint64_t o_seambits[UVOL_MAX_ATTR_DATA], o_eos[UVOL_MAX_ATTR_DATA], o_vos[UVOL_MAX_ATTR_DATA], o_ac2v[UVOL_MAX_ATTR_DATA],
             o_afirst[UVOL_MAX_ATTR_DATA], o_acnt[UVOL_MAX_ATTR_DATA];
    uint64_t o_seamcnt;                    // per chunk of corners: how many carry a seam bit (k_seam_count)
    uint64_t o_pcnt, o_pfirst;             // per-vertex point counts/offsets, dedup start corner
    uint64_t o_frec[UVOL_MAX_ATTR_DATA + 1], o_tstack[UVOL_MAX_ATTR_DATA + 1], o_fvis[UVOL_MAX_ATTR_DATA + 1];   // S, S, Z
    uint64_t o_corr[UVOL_MAX_ATTRS], o_par[UVOL_MAX_ATTRS], o_auxbits[UVOL_MAX_ATTRS];      // S.  Values are reconstructed IN PLACE over the corrections.
    uint32_t corr_cap[UVOL_MAX_ATTRS];     // capacity of o_corr[j] in symbols
    uint64_t o_tags[UVOL_MAX_ATTRS];       // S: decoded tags (int32 per entry) of attributes coded with the TAGGED scheme
    uint32_t table_cap[UVOL_MAX_ATTR_DATA + 1];   // capacity in entries of table t (attribute vertices); [0] = encoded vertices + splits (exact bound)
    // ---- count-sized arenas S2 / Z2 (zeroed) and the attribute part of the output arena: byte offsets filled ON THE DEVICE by
    // k_plan2 once the connectivity kernels have produced the counts (and recomputed by the host from the final counts)
    uint64_t o_p2c;                                                     // S2: point -> corner
    uint64_t o_d2c[UVOL_MAX_ATTR_DATA + 1], o_v2d[UVOL_MAX_ATTR_DATA + 1];   // S2, Z2
    // ---- outputs (byte offsets into the output arena): out_index header-sized (host), out_attr count-sized (device)
    uint64_t out_index, out_attr[4];
};

// Per-face traversal record of one corner table (base or attribute): built element-parallel (k_face_records),
// consumed by the speculative depth-first traversal as one 32-byte load per lane.
//   v[k]  vertex id of corner 3f+k in this table          o[k]  opposite corner of 3f+k (-1: boundary / seam)
//   meta  bits 0-1  corner through which the walk enters f from face f-1 (3: none)      bits 2-3  same from face f+1
//         bits 4-8  duplicate distance of the f-1 entry's tip vertex                      bits 9-13 same for the f+1 entry
//         bits 14-16 vertex of corner k lies on the table's boundary
struct FaceRec { int32_t v[3]; int32_t o[3]; uint32_t meta; uint32_t pad; };

// batch-level result of the device planner (lives behind the DracoCounts array)
// The output arena is region-major: [index buffers of all frames][positions of all frames][normals][colours][uvs], so a finished
// attribute of the whole batch is one contiguous span that can start its way to the host while later stages still run.
struct DracoBatchPlan { uint64_t s2_need, z2_need, out_need; uint64_t slot_base[4], slot_bytes[4]; uint32_t overflow, pad; };

// per-frame state written by the kernels (counts the host reads back once per batch)
struct DracoCounts {
    int32_t status;
    uint32_t num_vertex_slots;    // V (incl. isolated tail)
    uint32_t num_points;
    uint32_t attr_vertices[UVOL_MAX_ATTR_DATA];
    uint32_t entries[UVOL_MAX_ATTR_DATA + 1];   // per traversal table: [0] base, [1+i] attribute data i (written by the traversal)
    uint32_t expected[UVOL_MAX_ATTR_DATA + 1];  // entry counts known after phase 1 (valid vertices / attribute vertices)
    uint32_t rans_early[UVOL_MAX_ATTRS];        // symbols the early (self-terminating) attribute runs produced
    uint32_t dbg[4];
};

// ---- KTX2 / BasisLZ ------------------------------------------------------------------------
#define UVOL_HUFF_FAST_BITS 10
struct HuffTable {            // canonical Huffman decode table built on device
    uint32_t fast[1 << UVOL_HUFF_FAST_BITS];   // (sym << 8) | len ; len==0 -> slow path
    uint32_t first_code[17], first_idx[17], count[17];
    uint32_t total, used, maxl;
    uint32_t sorted_off;      // u16 offset into the per-file sorted-symbol pool
};

struct Ktx2File {
    uint64_t file_off; uint32_t file_len; int32_t status;
    uint32_t width, height, layers, is_video, has_alpha, is_uastc;
    uint32_t bx, by;
    uint32_t endpoint_count, selector_count;
    uint32_t ep_off, ep_len, sel_off, sel_len, tab_off, tab_len;   // byte offsets inside the file
    uint32_t level_off;                                            // level-0 payload offset inside the file
    uint32_t first_slice;                                          // index into the batch slice table
    uint32_t dfd_transfer, dfd_flags;
    // scratch / outputs
    uint64_t o_endpoints, o_selectors, o_huff, o_sorted;           // u8x4[ec], u8x4[sc], HuffTable[4], u16 pool
    uint64_t o_rgba;                                               // output arena offset, layers * w * h * 4
    uint32_t hist_size, pad;
    // Zstd-supercompressed level (scheme 2, UASTC only): inflated on the host into the blob at file_off (level_off = 0)
    uint32_t zstd, z_src_off, z_src_len, z_len;
};
struct Ktx2Slice {
    uint32_t file; uint32_t layer;
    uint32_t data_off, data_len;       // inside the file (absolute)
    uint32_t is_alpha, pad;
    uint64_t o_pred, o_delta, o_sel, o_ep;   // per-block scratch: u8 pred, u16 delta symbol, u16 selector, u16 endpoint
};
