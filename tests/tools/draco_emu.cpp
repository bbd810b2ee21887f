// draco_emu.cpp -- HOST EMULATION of the geometry pipeline's per-unit logic (draco_core.h), used
// only by tests to validate the algorithms without a GPU.  It runs the same __host__ __device__
// functions the kernels call, in the same stage order as the launcher in draco_decode.cu, with
// serial loops standing in for grids.  It is NOT part of libuvol_b200.so and is never a fallback.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../universal-volumetric_b200/csrc/uvol_internal.h"
#include "../../universal-volumetric_b200/csrc/draco_core.h"
#include "../../universal-volumetric_b200/csrc/draco_plan.h"

int uvol_draco_parse(const uint8_t *data, size_t len, DracoFrame &f, std::vector<uint32_t> &aux);


// ---- host emulation of the speculative traversal (k_traverse): 32 lanes as arrays, the warp collectives as loops.
static int g_spec_check = 1; static long g_spec_steps[UVOL_MAX_ATTR_DATA + 1]; static long g_spec_faces = 0;
extern "C" void draco_emu_spec(int enable, long *steps5, long *faces) { g_spec_check = enable; if (steps5) for (int i = 0; i <= UVOL_MAX_ATTR_DATA; i++) steps5[i] = g_spec_steps[i]; if (faces) *faces = g_spec_faces; }
static void build_face_records(const TableView &tv, const int *lmc, int F, FaceRec *rec) {
    for (int f = 0; f < F; f++) face_record(f, tv, lmc, rec[f]);
    for (int f = 0; f < F; f++) {          // duplicate distances (k_face_dups)
        const uint32_t mu = face_entry_tip(f, F, 1, tv), md = face_entry_tip(f, F, -1, tv); uint32_t du = 0, dd = 0;
        if (mu != 0xffffffffu) for (int k = 1; k < 32; k++) if (face_entry_tip(f - k, F, 1, tv) == mu) { du = (uint32_t)k; break; }
        if (md != 0xffffffffu) for (int k = 1; k < 32; k++) if (face_entry_tip(f + k, F, -1, tv) == md) { dd = (uint32_t)k; break; }
        rec[f].meta |= (du << 4) | (dd << 9);
    }
}
// The speculative traversal of k_traverse, lane by lane.  Row pattern: the walk often advances in ROWS -- k faces along the face order, then a jump of D
// faces (a strip crossed sideways: two faces per ring of a UV sphere, D = faces per ring) -- so lane l guesses face
// f0 + (l / k) * D + (l % k) * dir.  A lane at the start of a row finds its entry corner from its own record (the corner whose opposite
// lies in the previous lane's face); everything else -- the exact validation of every transition -- is as in traverse_spec_emu.
// (k, D) is adopted when two consecutive steps ended with the same row length and the same jump, and dropped on the first mismatch.
static int traverse_spec_emu(const FaceRec *rec, int F, uint8_t *fvis, int *v2d1, int *d2c, int *stk, int max_entries, uint32_t *out_n, long *steps) {
    const int C = 3 * F; int n = 0, sp = 0, c = -1, fscan = 0, pdir = 1; *steps = 0;
    int pk = 32, pD = 0, lastk = 0, lastD = 0;
    for (;;) {
        if (c < 0) {
            bool scan = false;
            for (;;) { if (sp == 0) { scan = true; break; } c = stk[sp - 1]; if (c < 0 || fvis[c / 3]) { sp--; c = -1; continue; } break; }
            if (scan) {
                int nf = fscan; while (nf < F && fvis[nf]) nf++;
                if (nf >= F) break;
                fscan = nf; c = 3 * nf; stk[0] = c; sp = 1;
                const int vn = rec[nf].v[1], vp = rec[nf].v[2];
                if (!v2d1[vn]) { if (n >= max_entries) return UVOL_ERR_CORRUPT; v2d1[vn] = ++n; d2c[n - 1] = c + 1; }
                if (!v2d1[vp]) { if (n >= max_entries) return UVOL_ERR_CORRUPT; v2d1[vp] = ++n; d2c[n - 1] = c + 2; }
            }
        }
        if (c >= C) return UVOL_ERR_CORRUPT;
        ++*steps;
        const int f0 = c / 3, k0 = c - 3 * f0;
        TravLane L[32]; bool selfopen[32], vis[32]; int act[32], nx[32], face[32];
        for (int l = 0; l < 32; l++) {
            const int fi = f0 + (l / pk) * pD + (l % pk) * pdir; face[l] = fi;
            const bool inr = fi >= 0 && fi < F;
            const FaceRec z{}; const FaceRec &r = inr ? rec[fi] : z;
            int kf = l == 0 ? k0 : -1;
            if (l > 0 && (l % pk) == 0) {          // row start: entered from the previous lane's face
                kf = 3;
                if (inr) for (int k = 0; k < 3; k++) if (r.o[k] >= 0 && r.o[k] / 3 == face[l - 1]) { kf = k; break; }
            }
            L[l] = trav_lane(r.v[0], r.v[1], r.v[2], r.o[0], r.o[1], r.o[2], r.meta, fi, kf, pdir, inr && kf != 3);
            selfopen[l] = L[l].ci >= 0 && !fvis[fi];
        }
        for (int l = 0; l < 32; l++) {
            bool dup = false;
            if (pk == 32) dup = l > 0 && (L[l].v == L[0].v || (L[l].pd != 0 && (int)L[l].pd < l));          // static distances along the face order
            else for (int e = 0; e < l; e++) if (L[e].ci >= 0 && L[e].v == L[l].v) dup = true;                // row pattern: the lanes compare their tips
            vis[l] = L[l].ci >= 0 && (v2d1[L[l].v] != 0 || dup);
            bool fr = true, fl = true;
            if (L[l].ci >= 0) {
                if (L[l].rc >= 0) { const int rf = L[l].rc / 3; fr = fvis[rf] != 0; for (int e = 0; e <= l; e++) if (face[e] == rf) fr = true; }
                if (L[l].lc >= 0) { const int lf = L[l].lc / 3; fl = fvis[lf] != 0; for (int e = 0; e <= l; e++) if (face[e] == lf) fl = true; }
            }
            trav_decide(vis[l], L[l].ob, fr, fl, L[l].rc, L[l].lc, &act[l], &nx[l]);
        }
        int m = 31;
        for (int l = 0; l < 32; l++) { const bool trans = act[l] == 0 && nx[l] >= 0 && l < 31 && nx[l] == L[l + 1].ci && selfopen[l + 1]; if (!trans) { m = l; break; } }
        if (getenv("EMU_TRAV_SEQ")) {          // debug trace of a window of steps (development aid)
            static long cnt2 = 0; const long w0 = atol(getenv("EMU_TRAV_SEQ"));
            if (cnt2 >= w0 && cnt2 < w0 + 70) fprintf(stderr, "step f0=%d k0=%d pdir=%d pk=%d pD=%d m=%d act=%d next=(face %d corner %d)\n", f0, k0, pdir, pk, pD, m, act[m], nx[m] >= 0 ? nx[m] / 3 : -1, nx[m] >= 0 ? nx[m] % 3 : -1);
            cnt2++;
        }
        if (!selfopen[0]) return UVOL_ERR_CORRUPT;
        for (int l = 0; l <= m; l++) {
            fvis[face[l]] = 1;
            if (!vis[l]) { if (n >= max_entries) return UVOL_ERR_CORRUPT; v2d1[L[l].v] = n + 1; d2c[n] = L[l].ci; n++; }
        }
        const int fm = face[m];
        if (act[m] == 0) { c = nx[m]; if (c < 0) return UVOL_ERR_CORRUPT; }
        else if (act[m] == 1) { sp--; c = -1; }
        else { stk[sp - 1] = L[m].lc; stk[sp] = nx[m]; sp++; c = nx[m]; }
        // next step's guess (as in the kernel)
        int nk = 32, nD = 0;
        if (c >= 0) {
            const int nf = c / 3;
            if (nf == fm + 1) pdir = 1; else if (nf == fm - 1) pdir = -1;
            else {
                const int row0 = pk == 32 ? 0 : (m / pk) * pk, rowlen = m - row0 + 1, D = nf - face[row0];
                const bool usable = rowlen <= 16 && (D > 32 || D < -32);
                if (usable && ((rowlen == lastk && D == lastD) || (pk < 32 && rowlen == pk && D == pD))) { nk = rowlen; nD = D; }
                lastk = rowlen; lastD = D;
            }
        } else { lastk = 0; lastD = 0; }
        if (pk < 32 && m == 31) { nk = pk; nD = pD; }          // a full step on the pattern: keep it
        pk = nk; pD = nD;
    }
    *out_n = (uint32_t)n;
    return UVOL_OK;
}

static void emu_rabs_bits(RabsLane &r, uint8_t *o, uint32_t n, bool toggle) {      // eight bits per store, as k_rabs_lanes
    uint32_t last = 1;
    for (uint32_t k = 0; k < n; k += 8) {
        uint8_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int b = 0; b < 8; b++) if (k + b < n) { uint32_t bit = rabs_lane_bit(r, true); if (toggle) { if (!bit) last ^= 1u; bit = last; } w[b] = (uint8_t)bit; }
        memcpy(o + k, w, 8);
    }
}

extern "C" int draco_emu_decode(const uint8_t *data, size_t len, uint32_t *num_points, uint32_t *num_faces,
                                uint32_t **index, float **position, float **normal, float **uv) {
    std::vector<DracoFrame> frames(1); std::vector<uint32_t> aux;
    DracoFrame &f = frames[0]; memset(&f, 0, sizeof f);
    f.file_off = 0; f.file_len = (uint32_t)len;
    f.status = uvol_draco_parse(data, len, f, aux);
    if (f.status) return f.status;
    DracoPlan pl; draco_plan_phase1(frames, pl);
    std::vector<uint8_t> scratch(pl.scratch + 256), zs(pl.zscratch + 256, 0);
    uint8_t *S = scratch.data(), *Z = zs.data();
    const uint8_t *file = data; aux.push_back(0);
    const int F = (int)f.nf, nad = (int)f.nad;
    DracoCounts cnt; memset(&cnt, 0, sizeof cnt);
    // stage: context symbol runs
    for (int i = 0; i < 6; i++) {
        const RansStream &s = f.ctx[i]; if (!s.count) continue;
        std::vector<uint32_t> cum(s.alphabet + 1, 0); std::vector<uint16_t> bucket(257);
        for (uint32_t k = 0; k < s.alphabet; k++) cum[k + 1] = cum[k] + aux[s.prob_off + k];
        if (cum[s.alphabet] != (1u << s.pb)) return UVOL_ERR_CORRUPT;
        for (uint32_t b = 0; b < 256; b++) bucket[b] = (uint16_t)rans_bucket_symbol(cum.data(), s.alphabet, b << (s.pb - 8));
        RansTables t{cum.data(), bucket.data(), s.alphabet, s.pb};
        int rc = rans_decode_run(file + s.data_off, s.data_len, t, s.count, 0, S + f.o_ctxsym[i]); if (rc) return rc;
    }
    // stage: seam bits
    for (int i = 0; i < nad; i++) {
        RabsLane r; if (!rabs_lane_init(r, file, f.seams[i])) return UVOL_ERR_CORRUPT;
        emu_rabs_bits(r, S + f.o_seambits[i], (uint32_t)(3 * F / 2 + 1), false);
    }
    // stage: edgebreaker
    EbMem m; m.opp = (int *)(S + f.o_opp); m.c2v = (int *)(S + f.o_c2v); m.lmc = (int *)(S + f.o_lmc); m.val = (int *)(S + f.o_val);
    m.hole = S + f.o_hole; m.stack = (int *)(S + f.o_stack); m.skey = m.stack + f.nsym + 8; m.sval = m.skey + f.nts + 1; m.invalid = (int *)(S + f.o_invalid);
    for (int i = 0; i < 6; i++) m.ctxsym[i] = S + f.o_ctxsym[i];
    int rc = eb_decode_frame(f, file, aux.data(), m, &cnt.num_vertex_slots); if (rc) return rc;
    const int V = (int)cnt.num_vertex_slots;
    // stage: seams
    { int idx = 0;
      for (int c = 0; c < 3 * F; c++) {
        int o = m.opp[c];
        if (o < 0) { for (int i = 0; i < nad; i++) seam_mark(c, m.opp, m.c2v, Z + f.o_eos[i], Z + f.o_vos[i]); continue; }
        if (o / 3 < c / 3) continue;
        for (int i = 0; i < nad; i++) if ((S + f.o_seambits[i])[idx]) seam_mark(c, m.opp, m.c2v, Z + f.o_eos[i], Z + f.o_vos[i]);
        idx++;
      } }
    // stage: attribute vertex tables (count, scan, assign)
    int err = 0;
    for (int i = 0; i < nad; i++) {
        int *acnt = (int *)(S + f.o_acnt[i]), *afirst = (int *)(S + f.o_afirst[i]), *ac2v = (int *)(S + f.o_ac2v[i]);
        for (int v = 0; v < V; v++) acnt[v] = attr_vertex_fan(v, m.opp, m.lmc, Z + f.o_eos[i], Z + f.o_vos[i], afirst, ac2v, 0, 0, F, &err);
        int run = 0; for (int v = 0; v < V; v++) { int c = acnt[v]; acnt[v] = run; run += c; }
        cnt.attr_vertices[i] = (uint32_t)run;
        for (int v = 0; v < V; v++) attr_vertex_fan(v, m.opp, m.lmc, Z + f.o_eos[i], Z + f.o_vos[i], afirst, ac2v, acnt[v], 1, F, &err);
    }
    const uint8_t *vos[UVOL_MAX_ATTR_DATA]; const int *ac2v[UVOL_MAX_ATTR_DATA];
    for (int i = 0; i < nad; i++) { vos[i] = Z + f.o_vos[i]; ac2v[i] = (int *)(S + f.o_ac2v[i]); }
    int *pcnt = (int *)(S + f.o_pcnt), *pfirst = (int *)(S + f.o_pfirst);
    for (int v = 0; v < V; v++) pcnt[v] = point_fan(v, m.opp, m.c2v, m.lmc, m.hole, nad, vos, ac2v, pfirst, nullptr, nullptr, 0, 0, F, &err);
    { int run = 0; for (int v = 0; v < V; v++) { int c = pcnt[v]; pcnt[v] = run; run += c; } cnt.num_points = (uint32_t)run; }
    if (err) return UVOL_ERR_CORRUPT;
    // ---- phase 2 (count-sized arrays: the same planning function the device planner kernel runs)
    draco_plan_phase2(frames, &cnt, pl);
    std::vector<uint8_t> scratch2(pl.scratch2 + 256), zs2(pl.zscratch2 + 256, 0), outb(pl.out + 256);
    uint8_t *S2 = scratch2.data(), *Z2 = zs2.data(), *O = outb.data();
    const int P = (int)cnt.num_points;
    uint32_t *c2p = (uint32_t *)(O + f.out_index); int *p2c = (int *)(S2 + f.o_p2c);
    for (int v = 0; v < V; v++) point_fan(v, m.opp, m.c2v, m.lmc, m.hole, nad, vos, ac2v, pfirst, c2p, p2c, pcnt[v], 1, F, &err);
    // traversals
    TableView tv[UVOL_MAX_ATTR_DATA + 1];
    for (int t = 0; t <= nad; t++) {
        if (f.o_d2c[t] == UVOL_NONE) continue;
        tv[t] = (t == 0) ? TableView{m.opp, m.c2v, nullptr, nullptr, nullptr} : TableView{m.opp, m.c2v, Z + f.o_eos[t - 1], ac2v[t - 1], vos[t - 1]};
        const int maxe = (int)(t == 0 ? cnt.num_vertex_slots : cnt.attr_vertices[t - 1]);
        std::vector<uint8_t> fvis_t(F + 4, 0);
        rc = traverse_table(tv[t], m.lmc, F, fvis_t.data(), (int *)(Z2 + f.o_v2d[t]), (int *)(S2 + f.o_d2c[t]), (int *)(S + f.o_tstack[t]), maxe, &cnt.entries[t]);
        if (rc) return rc;
        if (g_spec_check) {          // the speculative 32-lane traversal of k_traverse, emulated lane by lane on the same records: must give the same order
            std::vector<FaceRec> rec(F + 2); build_face_records(tv[t], m.lmc, F, rec.data());
            std::vector<uint8_t> fv2(F + 16, 0); std::vector<int> v2d2(maxe + 4, 0), d2c2(maxe + 4, 0), stk2(F + 8);
            uint32_t n2 = 0; long steps = 0;
            rc = traverse_spec_emu(rec.data(), F, fv2.data(), v2d2.data(), d2c2.data(), stk2.data(), maxe, &n2, &steps);
            if (rc) return rc;
            if (n2 != cnt.entries[t] || memcmp(d2c2.data(), S2 + f.o_d2c[t], (size_t)n2 * 4) || memcmp(v2d2.data(), Z2 + f.o_v2d[t], (size_t)maxe * 4)) return -77;
            g_spec_steps[t] = steps; g_spec_faces = F;
        }
    }
    // attribute symbol runs + aux bits
    for (int j = 0; j < f.nattr; j++) {
        const DracoAttr &a = f.attr[j]; if (f.o_corr[j] == UVOL_NONE) continue;
        const RansStream &s = a.sym; const uint32_t n = cnt.entries[a.table + 1];
        const int positive = a.pred != -2 && (a.xform == 2 || a.xform == 3);
        if (n * (uint32_t)a.vnc > f.corr_cap[j]) return UVOL_ERR_FRAME_CAPACITY;
        std::vector<uint32_t> cum(s.alphabet + 1, 0); std::vector<uint16_t> bucket(257);
        for (uint32_t k = 0; k < s.alphabet; k++) cum[k + 1] = cum[k] + aux[s.prob_off + k];
        if (cum[s.alphabet] != (1u << s.pb)) return UVOL_ERR_CORRUPT;
        for (uint32_t b = 0; b < 256; b++) bucket[b] = (uint16_t)rans_bucket_symbol(cum.data(), s.alphabet, b << (s.pb - 8));
        RansTables t{cum.data(), bucket.data(), s.alphabet, s.pb};
        if (!a.tagged) { rc = rans_decode_run(file + s.data_off, s.data_len, t, n * a.vnc, positive ? 2 : 1, S + f.o_corr[j]); if (rc) return rc; }
        else {      // TAGGED scheme: tags (k_rans, raw symbols) then the bit fields (k_tagged_values)
            if (f.o_tags[j] == UVOL_NONE) return -78;
            uint32_t *tags = (uint32_t *)(S + f.o_tags[j]); int32_t *out = (int32_t *)(S + f.o_corr[j]);
            rc = rans_decode_run(file + s.data_off, s.data_len, t, n, 2, tags); if (rc) return rc;
            const uint8_t *bits = file + a.tag_bits_off; uint64_t off = 0; const uint64_t total = 8ull * a.tag_bits_len;
            for (uint32_t e = 0; e < n; e++) {
                const uint32_t tg = tags[e]; if (tg > 32 || off + (uint64_t)tg * a.vnc > total) return UVOL_ERR_CORRUPT;
                for (int c = 0; c < a.vnc; c++, off += tg) {
                    uint64_t x = 0; for (int k = 0; k < 5 && (off >> 3) + k < (uint64_t)a.tag_bits_len; k++) x |= (uint64_t)bits[(off >> 3) + k] << (8 * k);
                    const uint32_t v = tg == 0 ? 0u : (uint32_t)(x >> (off & 7)) & (tg == 32 ? 0xffffffffu : ((1u << tg) - 1u));
                    out[e * a.vnc + c] = positive ? (int32_t)v : ((v & 1u) ? -(int32_t)(v >> 1) - 1 : (int32_t)(v >> 1));
                }
            }
        }
        if (a.pred == 5 || a.pred == 6) {
            RabsLane r; if (!rabs_lane_init(r, file, a.aux_bits)) return UVOL_ERR_CORRUPT;
            if (a.pred == 5) { if ((uint32_t)a.num_orient > n) return UVOL_ERR_CORRUPT; emu_rabs_bits(r, S + f.o_auxbits[j], (uint32_t)a.num_orient, true); }
            else emu_rabs_bits(r, S + f.o_auxbits[j], n, false);
        }
    }
    // prediction reversal: position-like first, then uv / normal
    const DracoAttr &pa = f.attr[f.pos_attr];
    const int *pos_v2d1 = (int *)(Z2 + f.o_v2d[0]); const int32_t *posq = (int32_t *)(S + f.o_corr[f.pos_attr]);
    for (int pass = 0; pass < 2; pass++) for (int j = 0; j < f.nattr; j++) {
        const DracoAttr &a = f.attr[j]; if (f.o_corr[j] == UVOL_NONE) continue;
        const int t = a.table + 1, n = (int)cnt.entries[t];
        const int32_t *corr = (int32_t *)(S + f.o_corr[j]); int32_t *val = (int32_t *)(S + f.o_corr[j]);          // values reconstructed in place
        const int *d2c = (int *)(S2 + f.o_d2c[t]), *v2d1 = (int *)(Z2 + f.o_v2d[t]);
        if (pass == 0 && (a.pred == 0 || a.pred == 1 || a.pred == -2)) {
            if (a.pred == -2) continue;
            int *par = (int *)(S + f.o_par[j]);
            if (a.pred == 1) for (int p = 0; p < n; p++) parallelogram_parents(p, tv[t], d2c, v2d1, par + 4 * p);
            for (int k = 0; k < a.vnc; k++) predict_wrap_component(k, a.vnc, n, a.pred == 1, par, corr, val, a.wmin, a.wmax);
        } else if (pass == 1 && a.pred == 5) {
            UvPrep *prep = (UvPrep *)(S + f.o_par[j]);
            for (int p = 0; p < n; p++) uv_prepare(p, tv[t], d2c, v2d1, pos_v2d1, posq, prep[p]);
            rc = predict_uv_chain(n, prep, corr, val, S + f.o_auxbits[j], a.num_orient, a.wmin, a.wmax); if (rc) return rc;
        } else if (pass == 1 && a.pred == 6) {
            for (int p = 0; p < n; p++) normal_entry(p, tv[t], d2c, pos_v2d1, posq, corr, S + f.o_auxbits[j], a.wmin, val);
        }
    }
    (void)pa;
    // expansion
    *num_points = (uint32_t)P; *num_faces = (uint32_t)F;
    *index = (uint32_t *)malloc((size_t)F * 12); memcpy(*index, c2p, (size_t)F * 12);
    float **dst[4] = {position, normal, uv, nullptr}; *position = *normal = *uv = nullptr;
    for (int j = 0; j < f.nattr; j++) {
        const DracoAttr &a = f.attr[j]; if (a.out_slot < 0 || a.out_slot > 2) continue;
        const int t = a.table + 1;
        float *o = (float *)(O + f.out_attr[a.out_slot]);
        const int *voc = t == 0 ? m.c2v : ac2v[t - 1];
        for (int p = 0; p < P; p++) expand_point(p, p2c, voc, (int *)(Z2 + f.o_v2d[t]), a, (int32_t *)(S + f.o_corr[j]), o);
        *dst[a.out_slot] = (float *)malloc((size_t)P * a.nc * 4); memcpy(*dst[a.out_slot], o, (size_t)P * a.nc * 4);
    }
    return err ? UVOL_ERR_CORRUPT : UVOL_OK;
}
extern "C" void draco_emu_free(void *p) { free(p); }
