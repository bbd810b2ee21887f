// draco_plan.h -- host-side memory planner for a geometry batch (shared by the CUDA launcher and
// by the host-emulation harness used to debug the logic).
//
// Phase 1 sizes depend only on header fields (F, encoded vertices, symbols); phase 2 sizes need
// the counts the connectivity kernels produce (vertex slots, attribute vertices, points), which
// the launcher reads back once per batch.  All offsets are bytes, 128-byte aligned.
#pragma once
#include <vector>
#include "uvol_internal.h"

struct DracoPlan { uint64_t scratch = 0, zscratch = 0, scratch2 = 0, zscratch2 = 0, tscratch = 0, out = 0, out_index = 0; };   // out_index: bytes of the index region at the head of the output arena

#define UVOL_NONE (~0ull)
static inline uint64_t plan_take(uint64_t &cursor, uint64_t bytes) { uint64_t o = cursor; cursor = (cursor + bytes + 127) / 128 * 128; return o; }

static inline void draco_plan_phase1(std::vector<DracoFrame> &frames, DracoPlan &pl) {
    uint64_t s = 0, z = 0;
    for (auto &f : frames) {
        if (f.status) continue;
        const uint64_t F = f.nf, C = 3 * F, maxv = (uint64_t)f.nv_enc + f.nsplit + 4;
        f.o_opp = plan_take(s, C * 4); f.o_c2v = plan_take(s, C * 4);
        f.o_lmc = plan_take(s, maxv * 4); f.o_val = plan_take(s, maxv * 16); f.o_hole = plan_take(s, maxv);   // o_val: int valences (generic path) or 16 B vertex records (valence path)
        f.o_stack = plan_take(s, ((uint64_t)f.nsym + 8) * 4 + ((uint64_t)f.nts + 1) * 8);
        f.o_invalid = plan_take(s, ((uint64_t)f.nsplit + 8) * 4);
        for (int i = 0; i < 6; i++) f.o_ctxsym[i] = plan_take(s, (uint64_t)f.ctx[i].count + 16);   // read back to front in aligned 8-byte words
        for (uint32_t i = 0; i < f.nad; i++) {
            f.o_seambits[i] = plan_take(s, C / 2 + 8);
            f.o_ac2v[i] = plan_take(s, C * 4);
            f.o_acnt[i] = plan_take(s, (maxv + 1) * 4);
            f.o_afirst[i] = plan_take(s, maxv * 4);        // first corner of each vertex fan
            f.o_eos[i] = plan_take(z, C); f.o_vos[i] = plan_take(z, maxv);
        }
        f.o_seamcnt = plan_take(s, (C / 8192 + 2) * 4);      // SEAM_CHUNK corners per count
        f.o_pcnt = plan_take(s, (maxv + 1) * 4);
        f.o_pfirst = plan_take(s, maxv * 4);               // dedup start corner per vertex
        for (int j = 0; j < f.nattr; j++) {                // early attribute symbol runs: capacity from the largest possible entry count
            const DracoAttr &a = f.attr[j];
            if (a.out_slot < 0 && j != f.pos_attr) { f.o_corr_early[j] = UVOL_NONE; f.corr_early_cap[j] = 0; continue; }
            // attribute tables can have up to C vertices; real meshes stay far below 2 per base vertex, and a run that hits the
            // capacity is simply decoded again by count once the count is known (k_corr_settle / k_rans)
            const uint64_t cap = (a.table < 0 ? maxv : (C < 2 * maxv + 1024 ? C : 2 * maxv + 1024)) * (uint64_t)a.vnc;
            f.corr_early_cap[j] = (uint32_t)cap; f.o_corr_early[j] = plan_take(s, (cap + 4) * 4);
        }
    }
    pl.scratch = s; pl.zscratch = z;
}

// counts[i] must hold the values read back from the device for frame i.
static inline void draco_plan_phase2(std::vector<DracoFrame> &frames, const DracoCounts *counts, DracoPlan &pl) {
    uint64_t s = 0, z = 0, o = 0, tr = 0;      // tr: the traversal-record arena (dead once the traversal is done; shareable between windows)
    for (size_t i = 0; i < frames.size(); i++) {
        DracoFrame &f = frames[i]; const DracoCounts &c = counts[i];
        if (f.status || c.status) continue;
        const uint64_t F = f.nf, P = c.num_points;
        f.o_p2c = plan_take(s, (P + 1) * 4);               // point -> corner
        bool need[UVOL_MAX_ATTR_DATA + 1] = {true, false, false, false, false};
        for (int j = 0; j < f.nattr; j++) if (f.attr[j].out_slot >= 0 || j == f.pos_attr) need[f.attr[j].table + 1] = true;
        for (uint32_t t = 0; t <= f.nad; t++) {
            if (!need[t]) { f.o_d2c[t] = f.o_v2d[t] = f.o_frec[t] = f.o_tstack[t] = f.o_fvis[t] = UVOL_NONE; continue; }
            const uint64_t nv = (t == 0 ? c.num_vertex_slots : c.attr_vertices[t - 1]) + 4;
            f.o_d2c[t] = plan_take(s, nv * 4); f.o_tstack[t] = plan_take(tr, (F + 8) * 4);
            f.o_frec[t] = plan_take(tr, (3 * F + 4) * 16 + 2 * (F + 4) * 16);      // per-corner traversal records + per-face up / down entry records
            f.o_v2d[t] = plan_take(z, nv * 4);
            f.o_fvis[t] = plan_take(z, F + 16);             // visited-face bytes of the global-map traversal
        }
        for (int j = 0; j < f.nattr; j++) {
            const DracoAttr &a = f.attr[j];
            if (a.out_slot < 0 && j != f.pos_attr) { f.o_corr[j] = f.o_val_attr[j] = f.o_par[j] = f.o_auxbits[j] = UVOL_NONE; continue; }
            const uint64_t n = (a.table < 0 ? c.num_vertex_slots : c.attr_vertices[a.table]) + 4;
            f.o_corr[j] = plan_take(s, n * a.vnc * 4); f.o_val_attr[j] = plan_take(s, n * a.vnc * 4);
            f.o_par[j] = plan_take(s, n * (a.pred == 5 ? 40 : 16));
            f.o_auxbits[j] = plan_take(s, n + 8);
        }
    }
    // Output arena: the index buffers of all frames first (final as soon as the points are assigned, so their copy to the host can
    // start while the traversal and prediction stages still run), then the per-point attribute arrays.
    for (size_t i = 0; i < frames.size(); i++) { DracoFrame &f = frames[i]; if (!f.status && !counts[i].status) f.out_index = plan_take(o, (uint64_t)f.nf * 12); }
    pl.out_index = o;
    for (size_t i = 0; i < frames.size(); i++) {
        DracoFrame &f = frames[i]; if (f.status || counts[i].status) continue;
        for (int k = 0; k < 4; k++) f.out_attr[k] = UVOL_NONE;
        for (int j = 0; j < f.nattr; j++) if (f.attr[j].out_slot >= 0) f.out_attr[f.attr[j].out_slot] = plan_take(o, (uint64_t)counts[i].num_points * f.attr[j].nc * 4);
    }
    pl.scratch2 = s; pl.zscratch2 = z; pl.tscratch = tr; pl.out = o;
}

// o_frec / o_tstack come out of phase 2 relative to the traversal-record arena; the kernels address everything relative to the
// phase-2 scratch base, so the launcher adds (traversal arena base - phase-2 scratch base), modulo 2^64.
static inline void draco_plan_rebase_traversal(std::vector<DracoFrame> &frames, uint64_t delta) {
    for (auto &f : frames) for (uint32_t t = 0; t <= UVOL_MAX_ATTR_DATA; t++) if (f.o_frec[t] != UVOL_NONE) { f.o_frec[t] += delta; f.o_tstack[t] += delta; }
}
