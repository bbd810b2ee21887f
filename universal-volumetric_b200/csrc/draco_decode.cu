// draco_decode.cu -- sm_100a kernels + batch launcher for the V2 geometry path.
//
// Replaces DRACOLoader.decodeGeometry -> DRACOWorker 'decode' (src/lib/DRACOLoader.js:104-187,
// 433-457, 470-590) for a whole batch of .drc files.  Stage order (DESIGN.md "Geometry pipeline"):
//   phase 1  rans_ctx -> rabs_seams -> edgebreaker -> seams -> attr_fan(count|scan|assign) -> point_fan(count|scan)
//   (one 4-byte-per-field count readback per frame; exact-size phase-2 buffers are planned from it)
//   phase 2  point_fan(assign) -> traverse -> rans_attr + rabs_aux -> parents -> predict_wrap ->
//            uv_prepare -> predict_uv | normals -> expand
// Serial units (entropy runs, connectivity walk, traversal, prediction chains) get one warp each and
// rely on the batch for parallelism; everything else is element-parallel over corners / vertices /
// entries / points of all frames.  No tensor-core work exists on this path (HBM / latency bound).
#include <chrono>
#include <string.h>
#include "uvol_ctx.h"
#include "draco_core.h"
#include "draco_plan.h"

namespace {

struct Job { uint32_t frame; int32_t what; };

__device__ __forceinline__ bool frame_dead(const DracoFrame *frames, const DracoCounts *counts, uint32_t f) {
    return frames[f].status != 0 || counts[f].status != 0;
}
__device__ __forceinline__ void frame_fail(DracoCounts *counts, uint32_t f, int code) { counts[f].status = code; }

// ---------------------------------------------------------------------------------------------
// rANS symbol runs: one warp per run.  The warp builds the cumulative table and a 256-bucket
// first-symbol index in shared memory, then lane 0 walks the run (strictly serial state chain).
// what: 0..5 = valence context i (u8 out) ; 16+j = attribute j (int32 out, zig-zag unless the
// transform yields positive corrections).
__global__ void __launch_bounds__(32) k_rans(const DracoFrame *frames, DracoCounts *counts, const uint8_t *blob, const uint32_t *aux,
                                             uint8_t *scratch, uint8_t *scratch2, const Job *jobs, int njobs) {
    extern __shared__ uint32_t smem[];
    if ((int)blockIdx.x >= njobs) return;
    const Job jb = jobs[blockIdx.x];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame];
    const uint8_t *file = blob + f.file_off;
    RansStream s; void *out; int mode; uint32_t count;
    if (jb.what < 16) { s = f.ctx[jb.what]; out = scratch + f.o_ctxsym[jb.what]; mode = 0; count = s.count; }
    else {
        const DracoAttr &a = f.attr[jb.what - 16];
        s = a.sym; out = scratch2 + f.o_corr[jb.what - 16];
        mode = (a.pred != -2 && (a.xform == 2 || a.xform == 3)) ? 2 : 1;
        count = counts[jb.frame].entries[a.table + 1] * (uint32_t)a.vnc;
    }
    const uint32_t A = s.alphabet, lane = threadIdx.x;
    uint32_t *cum = smem; uint16_t *bucket = (uint16_t *)(smem + A + 1);
    const uint32_t *prob = aux + s.prob_off;
    uint32_t run = 0;
    for (uint32_t base = 0; base < A; base += 32) {
        const uint32_t p = (base + lane < A) ? prob[base + lane] : 0u;
        uint32_t inc = p;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += t; }
        if (base + lane < A) cum[base + lane] = run + inc - p;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) cum[A] = run;
    __syncwarp();
    if (run != (1u << s.pb)) { if (lane == 0) frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }
    for (uint32_t b = lane; b < 256; b += 32) bucket[b] = (uint16_t)rans_bucket_symbol(cum, A, b << (s.pb - 8));
    __syncwarp();
    if (lane == 0) {
        RansTables t{cum, bucket, A, s.pb};
        const int rc = rans_decode_run(file + s.data_off, s.data_len, t, count, mode, out);
        if (rc) frame_fail(counts, jb.frame, rc);
    }
}

// rABS bit runs: one warp per run (lane 0 walks).  what: 0..3 = seam bits of attribute data i
// (upper bound 3F/2+1 bits); 16+j = attribute j aux bits (TEX_COORDS orientations incl. the
// toggle decoding, or GEOMETRIC_NORMAL flip bits).
__global__ void __launch_bounds__(32) k_rabs(const DracoFrame *frames, DracoCounts *counts, const uint8_t *blob,
                                             uint8_t *scratch, uint8_t *scratch2, const Job *jobs, int njobs) {
    if ((int)blockIdx.x >= njobs || threadIdx.x != 0) return;
    const Job jb = jobs[blockIdx.x];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame];
    const uint8_t *file = blob + f.file_off;
    Rabs r;
    if (jb.what < 16) {
        if (!rabs_init(r, file, f.seams[jb.what])) { frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }
        uint8_t *o = scratch + f.o_seambits[jb.what]; const int n = (int)(3 * f.nf / 2 + 1);
        for (int k = 0; k < n; k++) o[k] = (uint8_t)rabs_bit(r);
    } else {
        const DracoAttr &a = f.attr[jb.what - 16];
        if (!rabs_init(r, file, a.aux_bits)) { frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }
        uint8_t *o = scratch2 + f.o_auxbits[jb.what - 16];
        const uint32_t n = counts[jb.frame].entries[a.table + 1];
        if (a.pred == 5) {
            if ((uint32_t)a.num_orient > n) { frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }
            int last = 1;
            for (int k = 0; k < a.num_orient; k++) { if (!rabs_bit(r)) last = !last; o[k] = (uint8_t)last; }
        } else for (uint32_t k = 0; k < n; k++) o[k] = (uint8_t)rabs_bit(r);
    }
}

// Edgebreaker connectivity: one warp per frame, lane 0 walks the symbol sequence.
__global__ void __launch_bounds__(32) k_edgebreaker(const DracoFrame *frames, DracoCounts *counts, const uint8_t *blob, const uint32_t *aux,
                                                    uint8_t *S, int nframes) {
    const uint32_t fi = blockIdx.x;
    if ((int)fi >= nframes || threadIdx.x != 0) return;
    if (frames[fi].status) { counts[fi].status = frames[fi].status; return; }
    if (counts[fi].status) return;
    const DracoFrame &f = frames[fi];
    EbMem m; m.opp = (int *)(S + f.o_opp); m.c2v = (int *)(S + f.o_c2v); m.lmc = (int *)(S + f.o_lmc); m.val = (int *)(S + f.o_val);
    m.hole = S + f.o_hole; m.stack = (int *)(S + f.o_stack); m.skey = m.stack + f.nsym + 8; m.sval = m.skey + f.nts + 1; m.invalid = (int *)(S + f.o_invalid);
    for (int i = 0; i < 6; i++) m.ctxsym[i] = S + f.o_ctxsym[i];
    uint32_t slots = 0;
    const int rc = eb_decode_frame(f, blob + f.file_off, aux, m, &slots);
    counts[fi].num_vertex_slots = slots;
    if (rc) frame_fail(counts, fi, rc);
}

// Attribute seams: one CTA per frame.  The k-th bit of each seam stream belongs to the k-th corner
// (in corner order) whose opposite face is not older than its own; a ballot-based block scan turns
// that into a parallel lookup.
__global__ void __launch_bounds__(256) k_seams(const DracoFrame *frames, DracoCounts *counts, uint8_t *S, uint8_t *Z, int nframes) {
    __shared__ int warp_tot[8]; __shared__ int carry;
    const uint32_t fi = blockIdx.x;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    const int C = 3 * (int)f.nf, nad = (int)f.nad, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (nad == 0) return;
    const int *opp = (const int *)(S + f.o_opp), *c2v = (const int *)(S + f.o_c2v);
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < C; base += 256) {
        const int c = base + tid; int o = -2, flag = 0;
        if (c < C) { o = opp[c]; flag = (o >= 0 && o / 3 >= c / 3); }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int pre = carry;
        for (int k = 0; k < w; k++) pre += warp_tot[k];
        const int idx = pre + __popc(bal & ((1u << lane) - 1u));
        if (c < C) {
            if (o == -1) { for (int i = 0; i < nad; i++) seam_mark(c, opp, c2v, Z + f.o_eos[i], Z + f.o_vos[i]); }
            else if (flag) { for (int i = 0; i < nad; i++) if ((S + f.o_seambits[i])[idx]) seam_mark(c, opp, c2v, Z + f.o_eos[i], Z + f.o_vos[i]); }
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int k = 0; k < 8; k++) t += warp_tot[k]; carry += t; }
        __syncthreads();
    }
}

// RecomputeVertices per base vertex (count pass / assign pass).  grid = (ceil(maxV/128), frames, attribute data)
template <int PASS>
__global__ void __launch_bounds__(128) k_attr_fan(const DracoFrame *frames, DracoCounts *counts, uint8_t *S, uint8_t *Z) {
    const uint32_t fi = blockIdx.y, i = blockIdx.z;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if (i >= f.nad) return;
    const int V = (int)counts[fi].num_vertex_slots, v = blockIdx.x * 128 + threadIdx.x;
    if (v >= V) return;
    int err = 0; int *acnt = (int *)(S + f.o_acnt[i]);
    const int n = attr_vertex_fan(v, (const int *)(S + f.o_opp), (const int *)(S + f.o_lmc), Z + f.o_eos[i], Z + f.o_vos[i],
                                  (int *)(S + f.o_afirst[i]), (int *)(S + f.o_ac2v[i]), PASS ? acnt[v] : 0, PASS, (int)f.nf, &err);
    if (!PASS) acnt[v] = n;
    if (err) frame_fail(counts, fi, UVOL_ERR_CORRUPT);
}

// In-place exclusive scan of a per-vertex count array, total -> counts.  grid = (frames, arrays), 1024 threads.
// which: 0..3 attribute data i (acnt -> attr_vertices[i]); 4 points (pcnt -> num_points).
__global__ void __launch_bounds__(1024) k_scan(const DracoFrame *frames, DracoCounts *counts, uint8_t *S, int which0) {
    __shared__ int wsum[32]; __shared__ int carry_s;
    const uint32_t fi = blockIdx.x, which = which0 + blockIdx.y;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if (which < 4 && which >= f.nad) return;
    int *a = (int *)(S + (which < 4 ? f.o_acnt[which] : f.o_pcnt));
    const int V = (int)counts[fi].num_vertex_slots, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < V; base += 4096) {
        const int i0 = base + tid * 4; int x[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { x[k] = (i0 + k < V) ? a[i0 + k] : 0; sum += x[k]; }
        int inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        int pre = carry_s;
        for (int k = 0; k < w; k++) pre += wsum[k];
        int run = pre + inc - sum;
#pragma unroll
        for (int k = 0; k < 4; k++) { if (i0 + k < V) a[i0 + k] = run; run += x[k]; }
        __syncthreads();
        if (tid == 1023) carry_s = pre + inc;
        __syncthreads();
    }
    if (tid == 0) { if (which < 4) counts[fi].attr_vertices[which] = (uint32_t)carry_s; else counts[fi].num_points = (uint32_t)carry_s; }
}

// AssignPointsToCorners per base vertex.  PASS 0 counts (phase 1), PASS 1 writes the index buffer and point->corner.
template <int PASS>
__global__ void __launch_bounds__(128) k_point_fan(const DracoFrame *frames, DracoCounts *counts, uint8_t *S, uint8_t *Z, uint8_t *S2, uint8_t *O) {
    const uint32_t fi = blockIdx.y;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    const int V = (int)counts[fi].num_vertex_slots, v = blockIdx.x * 128 + threadIdx.x;
    if (v >= V) return;
    const uint8_t *vos[UVOL_MAX_ATTR_DATA]; const int *ac2v[UVOL_MAX_ATTR_DATA];
    for (uint32_t i = 0; i < UVOL_MAX_ATTR_DATA; i++) { vos[i] = i < f.nad ? Z + f.o_vos[i] : nullptr; ac2v[i] = i < f.nad ? (const int *)(S + f.o_ac2v[i]) : nullptr; }
    int err = 0; int *pcnt = (int *)(S + f.o_pcnt);
    const int *opp = (const int *)(S + f.o_opp), *c2v = (const int *)(S + f.o_c2v), *lmc = (const int *)(S + f.o_lmc);
    if (f.nad == 0) {   // identity: point id == vertex id (A.3)
        if (!PASS) { pcnt[v] = lmc[v] >= 0 ? 1 : 0; }
        else if (lmc[v] >= 0) {
            uint32_t *c2p = (uint32_t *)(O + f.out_index); int *p2c = (int *)(S2 + f.o_p2c);
            p2c[v] = lmc[v];
            int c = lmc[v], c0 = c, left = 1, guard = 0;
            while (c >= 0 && ++guard < 3 * (int)f.nf) {
                c2p[c] = (uint32_t)v;
                int nx;
                if (left) { nx = b_swl(opp, c); if (nx < 0) { nx = b_swr(opp, c0); left = 0; } else if (nx == c0) nx = DINV; } else nx = b_swr(opp, c);
                c = nx;
            }
        }
        return;
    }
    const int n = point_fan(v, opp, c2v, lmc, S + f.o_hole, (int)f.nad, vos, ac2v, (int *)(S + f.o_pfirst),
                            PASS ? (uint32_t *)(O + f.out_index) : nullptr, PASS ? (int *)(S2 + f.o_p2c) : nullptr, PASS ? pcnt[v] : 0, PASS, (int)f.nf, &err);
    if (!PASS) pcnt[v] = n;
    if (err) frame_fail(counts, fi, UVOL_ERR_CORRUPT);
}

__device__ __forceinline__ TableView make_view(const DracoFrame &f, int t, const uint8_t *S, const uint8_t *Z) {
    TableView tv; tv.opp = (const int *)(S + f.o_opp); tv.c2v_base = (const int *)(S + f.o_c2v);
    if (t == 0) { tv.eos = nullptr; tv.ac2v = nullptr; tv.vos = nullptr; }
    else { tv.eos = Z + f.o_eos[t - 1]; tv.ac2v = (const int *)(S + f.o_ac2v[t - 1]); tv.vos = Z + f.o_vos[t - 1]; }
    return tv;
}

// Depth-first traversal: one warp per (frame, table), lane 0 walks.  what = table (0 base, 1+i attribute data i).
__global__ void __launch_bounds__(32) k_traverse(const DracoFrame *frames, DracoCounts *counts, const uint8_t *S, const uint8_t *Z,
                                                 uint8_t *S2, uint8_t *Z2, const Job *jobs, int njobs) {
    if ((int)blockIdx.x >= njobs || threadIdx.x != 0) return;
    const Job jb = jobs[blockIdx.x];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame]; const int t = jb.what;
    if (f.o_d2c[t] == UVOL_NONE) return;
    const TableView tv = make_view(f, t, S, Z);
    const int maxe = (int)(t == 0 ? counts[jb.frame].num_vertex_slots : counts[jb.frame].attr_vertices[t - 1]);
    uint32_t n = 0;
    const int rc = traverse_table(tv, (const int *)(S + f.o_lmc), (int)f.nf, Z2 + f.o_fvis[t], (int *)(Z2 + f.o_v2d[t]), (int *)(S2 + f.o_d2c[t]),
                                  (int *)(S2 + f.o_tstack[t]), maxe, &n);
    counts[jb.frame].entries[t] = n;
    if (rc) frame_fail(counts, jb.frame, rc);
}

// Parallelogram parents, element-parallel.  grid = (ceil(maxN/128), frames, attrs)
__global__ void __launch_bounds__(128) k_parents(const DracoFrame *frames, const DracoCounts *counts, const uint8_t *S, const uint8_t *Z, uint8_t *S2, const uint8_t *Z2) {
    const uint32_t fi = blockIdx.y, j = blockIdx.z;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if ((int)j >= f.nattr || f.o_corr[j] == UVOL_NONE || f.attr[j].pred != 1) return;
    const int t = f.attr[j].table + 1, n = (int)counts[fi].entries[t], p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n) return;
    const TableView tv = make_view(f, t, S, Z);
    parallelogram_parents(p, tv, (const int *)(S2 + f.o_d2c[t]), (const int *)(Z2 + f.o_v2d[t]), (int *)(S2 + f.o_par[j]) + 4 * p);
}

// DIFFERENCE / PARALLELOGRAM + WRAP chain (also the pass-through for "no prediction"): one warp per
// (frame, attribute), lane k owns component k.  what = attribute index.
__global__ void __launch_bounds__(32) k_predict_wrap(const DracoFrame *frames, const DracoCounts *counts, uint8_t *S2, const Job *jobs, int njobs) {
    if ((int)blockIdx.x >= njobs) return;
    const Job jb = jobs[blockIdx.x];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame]; const int j = jb.what; const DracoAttr &a = f.attr[j];
    const int n = (int)counts[jb.frame].entries[a.table + 1], k = threadIdx.x;
    const int32_t *corr = (const int32_t *)(S2 + f.o_corr[j]); int32_t *val = (int32_t *)(S2 + f.o_val_attr[j]);
    if (a.pred == -2) { for (int i = k; i < n * a.vnc; i += 32) val[i] = corr[i]; return; }
    if (k >= a.vnc) return;
    predict_wrap_component(k, a.vnc, n, a.pred == 1, (const int *)(S2 + f.o_par[j]), corr, val, a.wmin, a.wmax);
}

// TEX_COORDS_PORTABLE position-only terms, element-parallel.  grid = (ceil(maxN/128), frames, attrs)
__global__ void __launch_bounds__(128) k_uv_prepare(const DracoFrame *frames, const DracoCounts *counts, const uint8_t *S, const uint8_t *Z, uint8_t *S2, const uint8_t *Z2) {
    const uint32_t fi = blockIdx.y, j = blockIdx.z;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if ((int)j >= f.nattr || f.o_corr[j] == UVOL_NONE || f.attr[j].pred != 5) return;
    const int t = f.attr[j].table + 1, n = (int)counts[fi].entries[t], p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n) return;
    const TableView tv = make_view(f, t, S, Z);
    UvPrep q;
    uv_prepare(p, tv, (const int *)(S2 + f.o_d2c[t]), (const int *)(Z2 + f.o_v2d[t]), (const int *)(Z2 + f.o_v2d[0]),
               (const int32_t *)(S2 + f.o_val_attr[f.pos_attr]), q);
    ((UvPrep *)(S2 + f.o_par[j]))[p] = q;
}

// TEX_COORDS_PORTABLE chain: one warp per (frame, attribute), lane 0 walks.
__global__ void __launch_bounds__(32) k_predict_uv(const DracoFrame *frames, DracoCounts *counts, uint8_t *S2, const Job *jobs, int njobs) {
    if ((int)blockIdx.x >= njobs || threadIdx.x != 0) return;
    const Job jb = jobs[blockIdx.x];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame]; const int j = jb.what; const DracoAttr &a = f.attr[j];
    const int n = (int)counts[jb.frame].entries[a.table + 1];
    const int rc = predict_uv_chain(n, (const UvPrep *)(S2 + f.o_par[j]), (const int32_t *)(S2 + f.o_corr[j]), (int32_t *)(S2 + f.o_val_attr[j]),
                                    S2 + f.o_auxbits[j], a.num_orient, a.wmin, a.wmax);
    if (rc) frame_fail(counts, jb.frame, rc);
}

// GEOMETRIC_NORMAL, element-parallel (each entry depends only on finished positions).
__global__ void __launch_bounds__(128) k_normals(const DracoFrame *frames, const DracoCounts *counts, const uint8_t *S, const uint8_t *Z, uint8_t *S2, const uint8_t *Z2) {
    const uint32_t fi = blockIdx.y, j = blockIdx.z;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if ((int)j >= f.nattr || f.o_corr[j] == UVOL_NONE || f.attr[j].pred != 6) return;
    const int t = f.attr[j].table + 1, n = (int)counts[fi].entries[t], p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n) return;
    const TableView tv = make_view(f, t, S, Z);
    normal_entry(p, tv, (const int *)(S2 + f.o_d2c[t]), (const int *)(Z2 + f.o_v2d[0]), (const int32_t *)(S2 + f.o_val_attr[f.pos_attr]),
                 (const int32_t *)(S2 + f.o_corr[j]), S2 + f.o_auxbits[j], f.attr[j].wmin, (int32_t *)(S2 + f.o_val_attr[j]));
}

// Per-point expansion + dequantisation into the output arrays.  grid = (ceil(maxP/256), frames, attrs)
__global__ void __launch_bounds__(256) k_expand(const DracoFrame *frames, const DracoCounts *counts, const uint8_t *S, const uint8_t *S2, const uint8_t *Z2, uint8_t *O) {
    const uint32_t fi = blockIdx.y, j = blockIdx.z;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if ((int)j >= f.nattr || f.attr[j].out_slot < 0) return;
    const int P = (int)counts[fi].num_points, p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const DracoAttr &a = f.attr[j]; const int t = a.table + 1;
    const int *voc = t == 0 ? (const int *)(S + f.o_c2v) : (const int *)(S + f.o_ac2v[t - 1]);
    expand_point(p, (const int *)(S2 + f.o_p2c), voc, (const int *)(Z2 + f.o_v2d[t]), a, (const int32_t *)(S2 + f.o_val_attr[j]), (float *)(O + f.out_attr[a.out_slot]));
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

static const char *kGeoStages[] = {"h2d", "rans_ctx", "rabs_seams", "edgebreaker", "seams", "attr_tables", "point_count", "counts_readback",
                                   "point_assign", "traverse", "rans_attr", "rabs_aux", "parents", "predict_wrap", "uv_prepare", "predict_uv", "normals",
                                   "expand", "d2h"};
extern "C" const char *uvol_geo_stage_name(int i) { return (i >= 0 && i < 19) ? kGeoStages[i] : ""; }

// Host-side state of the batch currently resident on the device (kept so that the device pipeline
// can be re-run on HBM-resident inputs, uvol_replay_draco_batch).
struct GeoBatch {
    std::vector<DracoFrame> frames; std::vector<uint32_t> aux; std::vector<Job> jobs;
    int n = 0, j_ransA = 0, j_rabsA = 0, j_trav = 0, j_ransB = 0, j_rabsB = 0, j_wrap = 0, j_uv = 0, j_end = 0;
    uint32_t max_alpha_ctx = 1, max_alpha_attr = 1, maxnad = 0, maxV = 0, maxF = 0; int maxattr = 0;
    uint64_t blob_bytes = 0, bytes_in = 0; DracoPlan pl; double parse_ms = 0;
};
void uvol_geo_batch_free(GeoBatch *b) { delete b; }

static int draco_prepare(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n) {
    const double t_begin = now_ms();
    if (!ctx->geo) ctx->geo = new GeoBatch();
    GeoBatch &B = *ctx->geo;
    B.n = n; B.frames.assign((size_t)n, DracoFrame()); B.aux.clear(); B.aux.reserve((size_t)n * 2048); B.jobs.clear();
    B.max_alpha_ctx = B.max_alpha_attr = 1; B.maxnad = B.maxV = B.maxF = 0; B.maxattr = 0; B.bytes_in = 0;
    std::vector<DracoFrame> &frames = B.frames; std::vector<uint32_t> &aux = B.aux;
    uint64_t blob_bytes = 0;
    for (int i = 0; i < n; i++) {
        DracoFrame &f = frames[i]; memset(&f, 0, sizeof f);
        f.file_off = blob_bytes; f.file_len = (uint32_t)size[i]; B.bytes_in += size[i];
        blob_bytes = align_up(blob_bytes + size[i] + 8, 16);
        f.status = (data[i] && size[i] < (1ull << 31)) ? uvol_draco_parse(data[i], size[i], f, aux) : UVOL_ERR_ARG;
        if (f.status) continue;
        for (int k = 0; k < 6; k++) if (f.ctx[k].count && f.ctx[k].alphabet > B.max_alpha_ctx) B.max_alpha_ctx = f.ctx[k].alphabet;
        for (int j = 0; j < f.nattr; j++) {
            if (f.attr[j].sym.alphabet > 49000) { f.status = UVOL_ERR_UNSUPPORTED; break; }
            if ((f.attr[j].out_slot >= 0 || j == f.pos_attr) && f.attr[j].sym.alphabet > B.max_alpha_attr) B.max_alpha_attr = f.attr[j].sym.alphabet;
        }
    }
    aux.push_back(0);
    B.blob_bytes = blob_bytes;
    UVOL_CUDA(ctx, ctx->h_blob.reserve(blob_bytes + 64));
    for (int i = 0; i < n; i++) if (data[i] && size[i] < (1ull << 31)) memcpy((uint8_t *)ctx->h_blob.p + frames[i].file_off, data[i], size[i]);
    draco_plan_phase1(frames, B.pl);
    std::vector<Job> &jobs = B.jobs; jobs.reserve((size_t)n * 24);
    auto mark = [&]() { return (int)jobs.size(); };
    B.j_ransA = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int k = 0; k < 6; k++) if (frames[i].ctx[k].count) jobs.push_back({(uint32_t)i, k});
    B.j_rabsA = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (uint32_t k = 0; k < frames[i].nad; k++) jobs.push_back({(uint32_t)i, (int)k});
    B.j_trav = mark();
    for (int i = 0; i < n; i++) {
        const DracoFrame &f = frames[i]; if (f.status) continue;
        bool need[UVOL_MAX_ATTR_DATA + 1] = {true, false, false, false, false};
        for (int j = 0; j < f.nattr; j++) if (f.attr[j].out_slot >= 0 || j == f.pos_attr) need[f.attr[j].table + 1] = true;
        for (uint32_t t = 0; t <= f.nad; t++) if (need[t]) jobs.push_back({(uint32_t)i, (int)t});
        if (f.nad > B.maxnad) B.maxnad = f.nad;
        if (f.nv_enc + f.nsplit > B.maxV) B.maxV = f.nv_enc + f.nsplit;
        if (f.nf > B.maxF) B.maxF = f.nf;
        if (f.nattr > B.maxattr) B.maxattr = f.nattr;
    }
    B.j_ransB = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int j = 0; j < frames[i].nattr; j++) if (frames[i].attr[j].out_slot >= 0 || j == frames[i].pos_attr) jobs.push_back({(uint32_t)i, 16 + j});
    B.j_rabsB = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int j = 0; j < frames[i].nattr; j++) if ((frames[i].attr[j].out_slot >= 0) && (frames[i].attr[j].pred == 5 || frames[i].attr[j].pred == 6)) jobs.push_back({(uint32_t)i, 16 + j});
    B.j_wrap = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int j = 0; j < frames[i].nattr; j++) { const DracoAttr &a = frames[i].attr[j]; if ((a.out_slot >= 0 || j == frames[i].pos_attr) && (a.pred == -2 || a.pred == 0 || a.pred == 1)) jobs.push_back({(uint32_t)i, j}); }
    B.j_uv = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int j = 0; j < frames[i].nattr; j++) if (frames[i].attr[j].out_slot >= 0 && frames[i].attr[j].pred == 5) jobs.push_back({(uint32_t)i, j});
    B.j_end = mark();
    // device buffers + uploads
    cudaStream_t st = ctx->s0;
    UVOL_CUDA(ctx, ctx->d_blob.reserve(blob_bytes + 64));
    UVOL_CUDA(ctx, ctx->d_desc.reserve(sizeof(DracoFrame) * (size_t)n));
    UVOL_CUDA(ctx, ctx->d_aux.reserve(aux.size() * 4));
    UVOL_CUDA(ctx, ctx->d_counts.reserve(sizeof(DracoCounts) * (size_t)n));
    UVOL_CUDA(ctx, ctx->h_counts.reserve(sizeof(DracoCounts) * (size_t)n));
    UVOL_CUDA(ctx, ctx->d_jobs.reserve(sizeof(Job) * (jobs.size() + 1)));
    UVOL_CUDA(ctx, ctx->d_scratch.reserve(B.pl.scratch + 256));
    UVOL_CUDA(ctx, ctx->d_zscratch.reserve(B.pl.zscratch + 256));
    UVOL_CUDA(ctx, ctx->h_desc.reserve(sizeof(DracoFrame) * (size_t)n + aux.size() * 4 + sizeof(Job) * (jobs.size() + 1)));
    B.parse_ms = now_ms() - t_begin;
    if (ctx->profile) cudaEventRecord(ctx->ev[0], st);
    uint8_t *hd = (uint8_t *)ctx->h_desc.p;
    uint8_t *h_aux = hd + sizeof(DracoFrame) * (size_t)n; memcpy(h_aux, aux.data(), aux.size() * 4);
    uint8_t *h_jobs = h_aux + aux.size() * 4; memcpy(h_jobs, jobs.data(), sizeof(Job) * jobs.size());
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_blob.p, ctx->h_blob.p, blob_bytes, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_aux.p, h_aux, aux.size() * 4, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_jobs.p, h_jobs, sizeof(Job) * jobs.size(), cudaMemcpyHostToDevice, st));
    return UVOL_OK;
}

static int draco_run(uvol_ctx *ctx, int memory, uvol_geometry *out, bool fresh_upload) {
    GeoBatch &B = *ctx->geo; const int n = B.n;
    std::vector<DracoFrame> &frames = B.frames; DracoPlan &pl = B.pl;
    cudaStream_t st = ctx->s0;
    int ev = 1;
    auto stamp = [&]() { if (ctx->profile && ev < 32) cudaEventRecord(ctx->ev[ev++], st); };
    if (!fresh_upload && ctx->profile) cudaEventRecord(ctx->ev[0], st);
    uint8_t *hd = (uint8_t *)ctx->h_desc.p;
    draco_plan_phase1(frames, pl);      // restores the phase-1 view of the descriptors (idempotent)
    memcpy(hd, frames.data(), sizeof(DracoFrame) * (size_t)n);
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_desc.p, hd, sizeof(DracoFrame) * (size_t)n, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_counts.p, 0, sizeof(DracoCounts) * (size_t)n, st));
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_zscratch.p, 0, pl.zscratch + 256, st));
    stamp();
    const DracoFrame *dF = (const DracoFrame *)ctx->d_desc.p; DracoCounts *dC = (DracoCounts *)ctx->d_counts.p;
    const uint8_t *dBlob = (const uint8_t *)ctx->d_blob.p; const uint32_t *dAux = (const uint32_t *)ctx->d_aux.p;
    uint8_t *dS = (uint8_t *)ctx->d_scratch.p, *dZ = (uint8_t *)ctx->d_zscratch.p; const Job *dJ = (const Job *)ctx->d_jobs.p;
    uint32_t launches = 0;
    auto rans_smem = [](uint32_t alphabet) { return (size_t)(alphabet + 1) * 4 + 257 * 2 + 16; };
    {
        const size_t smA = rans_smem(B.max_alpha_ctx), smB = rans_smem(B.max_alpha_attr);
        const size_t smMax = smA > smB ? smA : smB;
        if (smMax > 48 * 1024) UVOL_CUDA(ctx, cudaFuncSetAttribute(k_rans, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smMax));
    }
    // ---- phase 1
    if (B.j_rabsA - B.j_ransA > 0) { k_rans<<<B.j_rabsA - B.j_ransA, 32, rans_smem(B.max_alpha_ctx), st>>>(dF, dC, dBlob, dAux, dS, nullptr, dJ + B.j_ransA, B.j_rabsA - B.j_ransA); launches++; }
    stamp();
    if (B.j_trav - B.j_rabsA > 0) { k_rabs<<<B.j_trav - B.j_rabsA, 32, 0, st>>>(dF, dC, dBlob, dS, nullptr, dJ + B.j_rabsA, B.j_trav - B.j_rabsA); launches++; }
    stamp();
    k_edgebreaker<<<n, 32, 0, st>>>(dF, dC, dBlob, dAux, dS, n); launches++;
    stamp();
    k_seams<<<n, 256, 0, st>>>(dF, dC, dS, dZ, n); launches++;
    stamp();
    const unsigned gv = (B.maxV + 127) / 128 > 0 ? (B.maxV + 127) / 128 : 1;
    if (B.maxnad) {
        k_attr_fan<0><<<dim3(gv, n, B.maxnad), 128, 0, st>>>(dF, dC, dS, dZ); launches++;
        k_scan<<<dim3(n, B.maxnad), 1024, 0, st>>>(dF, dC, dS, 0); launches++;
        k_attr_fan<1><<<dim3(gv, n, B.maxnad), 128, 0, st>>>(dF, dC, dS, dZ); launches++;
    }
    stamp();
    k_point_fan<0><<<dim3(gv, n), 128, 0, st>>>(dF, dC, dS, dZ, nullptr, nullptr); launches++;
    k_scan<<<dim3(n, 1), 1024, 0, st>>>(dF, dC, dS, 4); launches++;
    stamp();
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts.p, dC, sizeof(DracoCounts) * (size_t)n, cudaMemcpyDeviceToHost, st));
    UVOL_CUDA(ctx, cudaStreamSynchronize(st));
    const DracoCounts *hC = (const DracoCounts *)ctx->h_counts.p;
    draco_plan_phase2(frames, hC, pl);
    UVOL_CUDA(ctx, ctx->d_scratch2.reserve(pl.scratch2 + 256));
    UVOL_CUDA(ctx, ctx->d_zscratch2.reserve(pl.zscratch2 + 256));
    UVOL_CUDA(ctx, ctx->d_out_geo.reserve(pl.out + 256));
    memcpy(hd, frames.data(), sizeof(DracoFrame) * (size_t)n);
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_desc.p, hd, sizeof(DracoFrame) * (size_t)n, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_zscratch2.p, 0, pl.zscratch2 + 256, st));
    stamp();
    uint8_t *dS2 = (uint8_t *)ctx->d_scratch2.p, *dZ2 = (uint8_t *)ctx->d_zscratch2.p, *dO = (uint8_t *)ctx->d_out_geo.p;
    uint32_t maxP = 1, maxN = 1;
    for (int i = 0; i < n; i++) if (!frames[i].status && !hC[i].status) {
        if (hC[i].num_points > maxP) maxP = hC[i].num_points;
        if (hC[i].num_vertex_slots > maxN) maxN = hC[i].num_vertex_slots;
        for (uint32_t k = 0; k < frames[i].nad; k++) if (hC[i].attr_vertices[k] > maxN) maxN = hC[i].attr_vertices[k];
    }
    // ---- phase 2
    k_point_fan<1><<<dim3(gv, n), 128, 0, st>>>(dF, dC, dS, dZ, dS2, dO); launches++;
    stamp();
    if (B.j_ransB - B.j_trav > 0) { k_traverse<<<B.j_ransB - B.j_trav, 32, 0, st>>>(dF, dC, dS, dZ, dS2, dZ2, dJ + B.j_trav, B.j_ransB - B.j_trav); launches++; }
    stamp();
    if (B.j_rabsB - B.j_ransB > 0) { k_rans<<<B.j_rabsB - B.j_ransB, 32, rans_smem(B.max_alpha_attr), st>>>(dF, dC, dBlob, dAux, dS, dS2, dJ + B.j_ransB, B.j_rabsB - B.j_ransB); launches++; }
    stamp();
    if (B.j_wrap - B.j_rabsB > 0) { k_rabs<<<B.j_wrap - B.j_rabsB, 32, 0, st>>>(dF, dC, dBlob, dS, dS2, dJ + B.j_rabsB, B.j_wrap - B.j_rabsB); launches++; }
    stamp();
    const unsigned gn = (maxN + 127) / 128;
    k_parents<<<dim3(gn, n, B.maxattr), 128, 0, st>>>(dF, dC, dS, dZ, dS2, dZ2); launches++;
    stamp();
    if (B.j_uv - B.j_wrap > 0) { k_predict_wrap<<<B.j_uv - B.j_wrap, 32, 0, st>>>(dF, dC, dS2, dJ + B.j_wrap, B.j_uv - B.j_wrap); launches++; }
    stamp();
    k_uv_prepare<<<dim3(gn, n, B.maxattr), 128, 0, st>>>(dF, dC, dS, dZ, dS2, dZ2); launches++;
    stamp();
    if (B.j_end - B.j_uv > 0) { k_predict_uv<<<B.j_end - B.j_uv, 32, 0, st>>>(dF, dC, dS2, dJ + B.j_uv, B.j_end - B.j_uv); launches++; }
    stamp();
    k_normals<<<dim3(gn, n, B.maxattr), 128, 0, st>>>(dF, dC, dS, dZ, dS2, dZ2); launches++;
    stamp();
    k_expand<<<dim3((maxP + 255) / 256, n, B.maxattr), 256, 0, st>>>(dF, dC, dS, dS2, dZ2, dO); launches++;
    stamp();
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts.p, dC, sizeof(DracoCounts) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (memory == UVOL_MEM_HOST) {
        UVOL_CUDA(ctx, ctx->h_out.reserve(pl.out + 256));
        UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_out.p, dO, pl.out, cudaMemcpyDeviceToHost, st));
    }
    stamp();
    UVOL_CUDA(ctx, cudaStreamSynchronize(st));
    UVOL_CUDA(ctx, cudaGetLastError());
    // ---- results
    uint8_t *base = memory == UVOL_MEM_HOST ? (uint8_t *)ctx->h_out.p : dO;
    uint64_t bytes_out = 0;
    for (int i = 0; i < n; i++) {
        const DracoFrame &f = frames[i]; uvol_geometry &g = out[i];
        memset(&g, 0, sizeof g);
        g.status = f.status ? f.status : hC[i].status;
        if (g.status) continue;
        g.num_points = hC[i].num_points; g.num_faces = f.nf;
        g.index = (uint32_t *)(base + f.out_index); bytes_out += (uint64_t)f.nf * 12;
        for (int j = 0; j < f.nattr; j++) {
            const DracoAttr &a = f.attr[j]; if (a.out_slot < 0) continue;
            float *p = (float *)(base + f.out_attr[a.out_slot]); bytes_out += (uint64_t)g.num_points * a.nc * 4;
            if (a.out_slot == 0) g.position = p; else if (a.out_slot == 1) g.normal = p; else if (a.out_slot == 2) g.uv = p; else { g.color = p; g.color_components = (uint32_t)a.nc; }
        }
    }
    uvol_stats &s = ctx->stats;
    s.kernel_launches = launches; s.bytes_in = B.bytes_in; s.bytes_out = bytes_out;
    s.scratch_bytes = pl.scratch + pl.zscratch + pl.scratch2 + pl.zscratch2;
    if (ctx->profile) {
        s.num_stages = (uint32_t)(ev - 1);
        for (int k = 0; k + 1 < ev && k < 24; k++) cudaEventElapsedTime(&s.stage_ms[k], ctx->ev[k], ctx->ev[k + 1]);
        float tot = 0; cudaEventElapsedTime(&tot, ctx->ev[1], ctx->ev[ev - 2]); s.device_ms = tot;   // kernels only: after h2d, before d2h
        s.h2d_ms = s.stage_ms[0]; s.d2h_ms = s.stage_ms[ev - 2];
    }
    return UVOL_OK;
}

extern "C" int uvol_decode_draco_batch(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int memory, uvol_geometry *out) {
    if (!ctx || !out || n < 0 || (n > 0 && (!data || !size))) return UVOL_ERR_ARG;
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->stats, 0, sizeof ctx->stats);
    if (n == 0) { if (ctx->geo) ctx->geo->n = 0; return UVOL_OK; }
    const double t0 = now_ms();
    int rc = draco_prepare(ctx, data, size, n); if (rc) return rc;
    rc = draco_run(ctx, memory, out, true); if (rc) return rc;
    ctx->stats.host_parse_ms = ctx->geo->parse_ms; ctx->stats.total_ms = now_ms() - t0;
    return UVOL_OK;
}

// Re-runs the device pipeline on the batch that is still resident in HBM from the last
// uvol_decode_draco_batch call (no parse, no input upload).  Used to time the kernels alone.
extern "C" int uvol_replay_draco_batch(uvol_ctx *ctx, int memory, uvol_geometry *out, int n) {
    if (!ctx || !out || !ctx->geo || ctx->geo->n != n || n <= 0) return UVOL_ERR_ARG;
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->stats, 0, sizeof ctx->stats);
    const double t0 = now_ms();
    const int rc = draco_run(ctx, memory, out, false); if (rc) return rc;
    ctx->stats.total_ms = now_ms() - t0;
    return UVOL_OK;
}
