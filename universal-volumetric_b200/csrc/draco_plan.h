// draco_plan.h -- memory planner of a geometry batch (shared by the CUDA launcher, the device planner kernel and the
// host-emulation harness used to debug the logic).
//
// Two kinds of arrays:
//   * header-sized (faces, encoded vertices, symbols, capacities derived from them): planned by the HOST before anything
//     runs (draco_plan_phase1) into the arenas S (uninitialised) and Z (zeroed), plus the index part of the output arena;
//   * count-sized (points, attribute vertices -- known only once the connectivity kernels have run): planned ON THE DEVICE
//     by k_plan2 (draco_plan2_frame below, the very same function the host runs again on the final counts), into the arenas
//     S2 / Z2 and the attribute part of the output arena.  The host reserves those arenas from optimistic estimates; a batch
//     that outgrows them is run again with the exact sizes (DracoBatchPlan.overflow).  No host round trip sits in the pipeline.
// All offsets are bytes, 128-byte aligned.
//
// S layout per frame: [long-lived: corner table, attribute corner tables, corrections / values, aux bits]
//                     [union region, three tenants one after the other in time:
//                        A  connectivity temporaries (vertex records, stacks, context symbols, seam bits, fan counts)   ... until the points are assigned
//                        B  traversal records (32 B per face per table) + traversal stacks                            ... k_face_records .. k_traverse
//                        C  prediction parents (16 B per entry; 40 B for the UV predictor)                            ... k_parents .. k_predict_uv ]
#pragma once
#include <vector>
#include "uvol_internal.h"

#if defined(__CUDACC__)
#define UVOL_HD __host__ __device__ __forceinline__
#else
#define UVOL_HD static inline
#endif

struct DracoPlan {
    uint64_t scratch = 0, zscratch = 0, out_index = 0;       // header-sized, exact (out_index: bytes of the index region at the head of the output arena)
    uint64_t s2_est = 0, z2_est = 0, out_est = 0;             // count-sized arenas: optimistic estimates (out_est includes the index region)
    uint64_t scratch2 = 0, zscratch2 = 0, out = 0;            // count-sized arenas: exact, known after the run
    uint32_t cap_entries = 1, cap_points = 1;                 // grid sizes of the entry- / point-parallel kernels (they stride, so these are not limits)
    bool base_records_early = false;                          // every frame's base-table traversal records lie clear of the connectivity temporaries
};

#define UVOL_NONE (~0ull)
UVOL_HD uint64_t plan_take(uint64_t &cursor, uint64_t bytes) { uint64_t o = cursor; cursor = (cursor + bytes + 127) / 128 * 128; return o; }
UVOL_HD bool draco_attr_needed(const DracoFrame &f, int j) { return f.attr[j].out_slot >= 0 || j == f.pos_attr; }
UVOL_HD void draco_tables_needed(const DracoFrame &f, bool need[UVOL_MAX_ATTR_DATA + 1]) {
    need[0] = true;
    for (int t = 1; t <= UVOL_MAX_ATTR_DATA; t++) need[t] = false;
    for (int j = 0; j < f.nattr; j++) if (draco_attr_needed(f, j)) need[f.attr[j].table + 1] = true;
}

// cap_permille scales the optimistic attribute-table capacity (1000 = the default 2 * vertices + 1024); tests shrink it to force
// the re-plan path.
static inline void draco_plan_phase1(std::vector<DracoFrame> &frames, DracoPlan &pl, uint32_t cap_permille = 1000) {
    uint64_t s = 0, z = 0, o = 0, s2 = 0, z2 = 0, oa = 0;
    pl.cap_entries = pl.cap_points = 1; pl.base_records_early = true;
    for (auto &f : frames) if (!f.status) f.out_index = plan_take(o, (uint64_t)f.nf * 12);
    pl.out_index = o;
    for (auto &f : frames) {
        if (f.status) continue;
        const uint64_t F = f.nf, C = 3 * F, nvmax = (uint64_t)f.nv_enc + f.nsplit, maxv = nvmax + 4;
        bool need[UVOL_MAX_ATTR_DATA + 1]; draco_tables_needed(f, need);
        f.table_cap[0] = (uint32_t)nvmax;
        for (uint32_t t = 1; t <= UVOL_MAX_ATTR_DATA; t++) {
            uint64_t cap = (2 * maxv + 1024) * cap_permille / 1000;
            if (f.full_cap || cap > C) cap = C;
            f.table_cap[t] = (uint32_t)cap;
        }
        // ---- long-lived
        f.o_opp = plan_take(s, C * 4); f.o_c2v = plan_take(s, C * 4);
        f.o_lmc = plan_take(s, maxv * 4); f.o_hole = plan_take(s, maxv);
        for (uint32_t i = 0; i < f.nad; i++) f.o_ac2v[i] = plan_take(s, C * 4);
        f.o_seamcnt = plan_take(s, (C / 8192 + 2) * 4);      // SEAM_CHUNK corners per count
        for (int j = 0; j < UVOL_MAX_ATTRS; j++) { f.o_corr[j] = f.o_par[j] = f.o_auxbits[j] = f.o_tags[j] = UVOL_NONE; f.corr_cap[j] = 0; }
        for (int j = 0; j < f.nattr; j++) {
            const DracoAttr &a = f.attr[j];
            if (!draco_attr_needed(f, j)) continue;
            const uint64_t cap = (uint64_t)f.table_cap[a.table + 1] * (uint64_t)a.vnc;
            f.corr_cap[j] = (uint32_t)cap; f.o_corr[j] = plan_take(s, (cap + 4) * 4);
            if (a.pred == 5 || a.pred == 6) f.o_auxbits[j] = plan_take(s, (uint64_t)f.table_cap[a.table + 1] + 8);
            if (a.tagged) f.o_tags[j] = plan_take(s, ((uint64_t)f.table_cap[a.table + 1] + 4) * 4);
        }
        // ---- union region
        const uint64_t U = s;
        uint64_t a = U, b = U, c = U;
        f.o_val = plan_take(a, maxv * 16);                    // int valences (generic path) or 16 B vertex records (valence path)
        f.o_stack = plan_take(a, ((uint64_t)f.nsym + 8) * 4 + ((uint64_t)f.nts + 1) * 8);
        f.o_invalid = plan_take(a, ((uint64_t)f.nsplit + 8) * 4);
        for (int i = 0; i < 6; i++) f.o_ctxsym[i] = plan_take(a, (uint64_t)f.ctx[i].count + 16);   // read back to front in aligned 8-byte words
        for (uint32_t i = 0; i < f.nad; i++) {
            f.o_seambits[i] = plan_take(a, C / 2 + 8);
            f.o_acnt[i] = plan_take(a, (maxv + 1) * 4);
            f.o_afirst[i] = plan_take(a, maxv * 4);            // first corner of each vertex fan
        }
        f.o_pcnt = plan_take(a, (maxv + 1) * 4);
        f.o_pfirst = plan_take(a, maxv * 4);                  // dedup start corner per vertex
        // attribute tables first, the base table LAST: its records then start behind tenant A whenever an attribute table exists, so
        // they can be built as soon as the connectivity is decoded (next to the seam / attribute-table / point stages)
        for (int t = UVOL_MAX_ATTR_DATA; t >= 0; t--) {
            if ((uint32_t)t > f.nad || !need[t]) { f.o_frec[t] = f.o_tstack[t] = f.o_fvis[t] = UVOL_NONE; continue; }
            f.o_frec[t] = plan_take(b, (F + 2) * sizeof(FaceRec));
            f.o_tstack[t] = plan_take(b, (F + 8) * 4);
        }
        if (f.o_frec[0] < a) pl.base_records_early = false;
        for (int j = 0; j < f.nattr; j++) {
            const DracoAttr &at = f.attr[j];
            if (!draco_attr_needed(f, j) || !(at.pred == 0 || at.pred == 1 || at.pred == 5)) continue;
            f.o_par[j] = plan_take(c, ((uint64_t)f.table_cap[at.table + 1] + 4) * (at.pred == 5 ? 40 : 16));
        }
        s = a > b ? a : b; if (c > s) s = c;
        // ---- zeroed
        for (uint32_t i = 0; i < f.nad; i++) { f.o_eos[i] = plan_take(z, C); f.o_vos[i] = plan_take(z, maxv); }
        for (uint32_t t = 0; t <= f.nad; t++) if (need[t]) f.o_fvis[t] = plan_take(z, F + 16);      // visited-face bytes of the global-map traversal
        // ---- optimistic estimates of the count-sized arrays: about 1.3 points / attribute vertices per encoded vertex
        uint64_t est = maxv * 13 / 10 + 256, pest = est < C ? est : C;
        s2 += ((pest + 1) * 4 + 127) / 128 * 128;
        for (uint32_t t = 0; t <= f.nad; t++) if (need[t]) {
            uint64_t e = t == 0 ? maxv : (est < f.table_cap[t] ? est : (uint64_t)f.table_cap[t] + 4);
            s2 += (e * 4 + 127) / 128 * 128; z2 += (e * 4 + 127) / 128 * 128;
        }
        for (int j = 0; j < f.nattr; j++) if (f.attr[j].out_slot >= 0) oa += (pest * f.attr[j].nc * 4 + 127) / 128 * 128;
        if (pest > pl.cap_points) pl.cap_points = (uint32_t)pest;      // grid bounds only: the entry / point kernels stride over any excess
        if (pest > pl.cap_entries) pl.cap_entries = (uint32_t)pest;
        if (maxv > pl.cap_entries) pl.cap_entries = (uint32_t)maxv;
    }
    pl.scratch = s; pl.zscratch = z; pl.s2_est = s2; pl.z2_est = z2; pl.out_est = o + oa;
}

// The count-sized arrays of one frame, laid out from the cursors (S2, Z2, and one per output slot: the output arena is region-major,
// see DracoBatchPlan).  assign = false only advances the cursors (sizes are independent of the start because every array is padded
// to 128 bytes).
struct Plan2Cursor { uint64_t s, z, o[4]; };
#define UVOL_SLOT_ORDER {0, 1, 3, 2}          // regions in the arena: position, normal, colour, uv (uv is finished last)
UVOL_HD void draco_plan2_frame(DracoFrame &f, const DracoCounts &c, Plan2Cursor &cur, bool assign) {
    if (f.status || c.status) return;
    const uint64_t P = c.num_points;
    uint64_t off = plan_take(cur.s, (P + 1) * 4);
    if (assign) f.o_p2c = off;
    bool need[UVOL_MAX_ATTR_DATA + 1]; draco_tables_needed(f, need);
    for (uint32_t t = 0; t <= UVOL_MAX_ATTR_DATA; t++) {
        if (t > f.nad || !need[t]) { if (assign) f.o_d2c[t] = f.o_v2d[t] = UVOL_NONE; continue; }
        const uint64_t nv = (t == 0 ? c.num_vertex_slots : c.attr_vertices[t - 1]) + 4;
        off = plan_take(cur.s, nv * 4); if (assign) f.o_d2c[t] = off;
        off = plan_take(cur.z, nv * 4); if (assign) f.o_v2d[t] = off;
    }
    if (assign) for (int k = 0; k < 4; k++) f.out_attr[k] = UVOL_NONE;
    for (int j = 0; j < f.nattr; j++) if (f.attr[j].out_slot >= 0) {
        const int k = f.attr[j].out_slot;
        off = plan_take(cur.o[k], P * (uint64_t)f.attr[j].nc * 4);
        if (assign) f.out_attr[k] = off;
    }
}
// Region bases from the per-slot totals (both on the device and on the host).
UVOL_HD void draco_plan2_regions(uint64_t out_index_bytes, const uint64_t total[4], uint64_t base[4], uint64_t *out_need) {
    const int order[4] = UVOL_SLOT_ORDER; uint64_t cur = out_index_bytes;
    for (int i = 0; i < 4; i++) { base[order[i]] = cur; cur += total[order[i]]; }
    *out_need = cur;
}

// Host: the exact layout from the final counts (identical to what k_plan2 wrote into the device descriptors).
static inline void draco_plan_phase2(std::vector<DracoFrame> &frames, const DracoCounts *counts, DracoPlan &pl) {
    Plan2Cursor tot{0, 0, {0, 0, 0, 0}};
    for (size_t i = 0; i < frames.size(); i++) draco_plan2_frame(frames[i], counts[i], tot, false);
    Plan2Cursor cur{0, 0, {0, 0, 0, 0}};
    draco_plan2_regions(pl.out_index, tot.o, cur.o, &pl.out);
    for (size_t i = 0; i < frames.size(); i++) draco_plan2_frame(frames[i], counts[i], cur, true);
    pl.scratch2 = cur.s; pl.zscratch2 = cur.z;
}
