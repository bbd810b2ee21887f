// draco_decode.cu -- sm_100a kernels + batch launcher for the V2 geometry path.
//
// Replaces DRACOLoader.decodeGeometry -> DRACOWorker 'decode' (src/lib/DRACOLoader.js:104-187,
// 433-457, 470-590) for a whole batch of .drc files.  Stage order (DESIGN.md "Geometry pipeline"):
//   phase 1  rans_ctx -> rabs_seams -> edgebreaker -> seams -> attr_fan(count|scan|assign) -> point_fan(count|scan)
//   plan2    the count-sized arrays (points, attribute vertices) are laid out ON THE DEVICE from the counts phase 1 produced
//            (k_plan2; no host round trip: the host reads the counts once, after the last kernel)
//   phase 2  point_fan(assign) -> face_records -> traverse -> rans_attr + rabs_aux -> parents -> predict_wrap ->
//            normals | uv_prepare -> predict_uv -> expand
// Serial units (entropy runs, connectivity walk, traversal, prediction chains) get one warp each and
// rely on the batch for parallelism; everything else is element-parallel over corners / vertices /
// entries / points of all frames.  No tensor-core work exists on this path (HBM / latency bound).
#include <chrono>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <thread>
#include "uvol_ctx.h"
#include "draco_core.h"
#include "draco_plan.h"

namespace {

struct Job { uint32_t frame; int32_t what; };
// UvPrep for the serial chain (40 bytes).  rcp > 0: pn2 / dot / ns are exact doubles and every product of the chain stays
// below 2^52 (fast path, all arithmetic exact in fp64); rcp == 0: |PN|^2 == 0; rcp < 0: the three fields hold the
// long long bit patterns (64-bit integer path).
struct UvPrepD { int32_t nd, pd; double pn2, dot, ns, rcp; };

__device__ __forceinline__ bool frame_dead(const DracoFrame *frames, const DracoCounts *counts, uint32_t f) {
    return frames[f].status != 0 || counts[f].status != 0;
}
__device__ __forceinline__ void frame_fail(DracoCounts *counts, uint32_t f, int code) { counts[f].status = code; }

// ---------------------------------------------------------------------------------------------
// rANS symbol runs: one warp per run.  The warp builds the cumulative table and a 256-bucket
// first-symbol index in shared memory, then lane 0 walks the run (strictly serial state chain).
// what: 0..5 = valence context i (u8 out) ; 16+j = attribute j (int32 out, zig-zag unless the
// transform yields positive corrections).
// All serial kernels give one unit of work to one warp, SERIAL_WARPS warps per block.  One warp per block keeps
// the shared-memory footprint of a block small, so the blocks of the concurrently running stages (geometry main
// stream, entropy side stream, texture stream) pack onto the SMs without forcing a second wave.
#ifndef SERIAL_WARPS
#define SERIAL_WARPS 1
#endif
// The serial rANS walk of one run (lane 0).  Everything on the dependent chain is kept to one shared-memory load per
// symbol: a 1024-entry index keyed by the top bits of the slot holds the first symbol reaching into that bucket together
// with its {start, frequency}; only when the slot lies beyond that symbol does the walk scan on through the compact table
// of the symbols that have a non-zero probability (kidx = that symbol's rank).  WIDE = alphabet > 4096 (symbol id split
// over both words).  EARLY = run until the coder is back in its initial state with all bytes consumed (at most `count`
// symbols): while no byte is left every encoder step strictly increases the state, so the initial state can only be met at the
// true end of the run -- this lets the attribute runs start before the connectivity has produced their symbol counts.
struct RansRun { const uint4 *cs; const uint2 *lut; const uint16_t *kidx; const uint32_t *wp; uint32_t w, wn, bi; int left; uint32_t st, pb, shift; };
template <int MODE, bool WIDE, bool EARLY>
__device__ __forceinline__ uint32_t rans_walk(RansRun r, uint32_t count, void *out) {
    const uint4 *cs = r.cs; const uint2 *lut = r.lut; const uint32_t *wp = r.wp;
    uint32_t w = r.w, wn = r.wn, bi = r.bi, st = r.st; int left = r.left;
    const uint32_t pb = r.pb, shift = r.shift, prec = 1u << pb, lbase = prec * 4u;
    uint8_t *o8 = (uint8_t *)out; int32_t *o32 = (int32_t *)out;
    uint32_t i = 0;
#define RANS_SYMBOL() do { \
        const uint32_t q = st >> pb, rem = st & (prec - 1); \
        const uint2 e = lut[rem >> shift]; \
        uint32_t sy = WIDE ? (e.x >> 20) | ((e.y >> 21) << 12) : e.x >> 20, start = e.x & 0xfffffu, freq = WIDE ? e.y & 0x1fffffu : e.y; \
        if (rem >= start + freq) { uint32_t k = r.kidx[rem >> shift]; do { const uint4 t = cs[++k]; start = t.x; freq = t.y; sy = t.z; } while (rem >= start + freq); } \
        st = q * freq + rem - start; \
        if (MODE == 0) o8[i] = (uint8_t)sy; \
        else if (MODE == 1) o32[i] = (sy & 1) ? -(int32_t)(sy >> 1) - 1 : (int32_t)(sy >> 1); \
        else o32[i] = (int32_t)sy; } while (0)
#pragma unroll 4
    for (; i < count && (!EARLY || left > 0); i++) {
        if (st < lbase) {
            while (st < lbase && left > 0) {
                st = st * 256u + (__byte_perm(w, 0u, bi | 0x4440u));
                left--;
                if (bi == 0) { w = wn; --wp; wn = wp[-1]; bi = 3; } else bi--;
            }
            if (EARLY && left == 0 && st == lbase) return i;
        }
        RANS_SYMBOL();
    }
    if (EARLY) for (; i < count && st != lbase; i++) RANS_SYMBOL();          // all bytes consumed: run down to the initial state
#undef RANS_SYMBOL
    return i;
}

// lut_bits: size of the slot index (8 B + 2 B per entry) -- 10 for the attribute runs (alphabets of hundreds of symbols), 6 for the
// connectivity context runs (five symbols): their tables then take under 1 KB, four runs share a block, and all six runs of every
// frame of a 1000-frame batch are resident at once (with the 10-bit index they needed two waves and sat at the head of the critical path).
#define RANS_LUT_BITS_MAX 10
__global__ void __launch_bounds__(128) k_rans(const DracoFrame *frames, DracoCounts *counts, const uint8_t *blob, const uint32_t *aux,
                                             uint8_t *scratch, const Job *jobs, int njobs, int smem_words_per_warp, int early, int lut_bits) {
    extern __shared__ uint32_t smem_all[];
    const int ji = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ji >= njobs) return;
    uint32_t *smem = smem_all + (size_t)(threadIdx.x >> 5) * smem_words_per_warp;
    const Job jb = jobs[ji];
    const uint32_t lane = threadIdx.x & 31;
    if (early) { if (frames[jb.frame].status) return; }
    else if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame];
    const uint8_t *file = blob + f.file_off;
    RansStream s; void *out; int mode; uint32_t count;
    if (jb.what < 16) { s = f.ctx[jb.what]; out = scratch + f.o_ctxsym[jb.what]; mode = 0; count = s.count; }
    else {
        const int j = jb.what - 16; const DracoAttr &a = f.attr[j];
        s = a.sym;
        // TAGGED scheme: the run holds one bit-length tag per entry (raw symbols into o_tags; k_tagged_values reads the components)
        const bool tg = a.tagged != 0;
        mode = (tg || (a.pred != -2 && (a.xform == 2 || a.xform == 3))) ? 2 : 1;
        out = scratch + (tg ? f.o_tags[j] : f.o_corr[j]);
        const uint32_t cap = tg ? f.table_cap[a.table + 1] : f.corr_cap[j];
        if (early) { count = cap; if (lane == 0) counts[jb.frame].rans_early[j] = 0xffffffffu; }
        else {
            count = counts[jb.frame].expected[a.table + 1] * (tg ? 1u : (uint32_t)a.vnc);
            if (counts[jb.frame].rans_early[j] == count) return;               // the early run already produced exactly these symbols, in place
            if (count > cap) { if (lane == 0) frame_fail(counts, jb.frame, UVOL_ERR_FRAME_CAPACITY); return; }
        }
    }
    // Tables in shared memory (built by the whole warp): cs[k] = {first slot, frequency, symbol} of the k-th symbol with a
    // non-zero probability (cs[nnz].x = total), lut[b] = {first slot | symbol, frequency} and kidx[b] = rank of the first
    // symbol whose range reaches into bucket b.
    const uint32_t A = s.alphabet, pb = s.pb;
    const uint32_t lb = pb < (uint32_t)lut_bits ? pb : (uint32_t)lut_bits;
    if (A > (1u << 18) || pb > 20u) { if (lane == 0) frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }      // the RAW scheme allows 18-bit symbols, 20-bit precision
    uint2 *lut = (uint2 *)smem; uint16_t *kidx = (uint16_t *)(lut + (1u << lut_bits)); uint4 *cs = (uint4 *)(kidx + (1u << lut_bits));
    const uint32_t *prob = aux + s.prob_off;
    uint32_t run = 0, nk = 0;
    for (uint32_t base = 0; base < A; base += 32) {
        const uint32_t p = (base + lane < A) ? prob[base + lane] : 0u;
        uint32_t inc = p;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += t; }
        const unsigned nzm = __ballot_sync(0xffffffffu, p != 0);
        if (p != 0) { const uint32_t k = nk + __popc(nzm & ((1u << lane) - 1u)); if (k < s.nnz) cs[k] = make_uint4(run + inc - p, p, base + lane, 0u); }
        run += __shfl_sync(0xffffffffu, inc, 31); nk += __popc(nzm);
    }
    if (lane == 0) cs[s.nnz] = make_uint4(run, 0u, A, 0u);
    __syncwarp();
    if (run != (1u << pb) || nk != s.nnz) { if (lane == 0) frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }
    for (uint32_t b = lane; b < (1u << lb); b += 32) {          // first k with cs[k+1].start > b << (pb - lb)
        const uint32_t target = b << (pb - lb);
        uint32_t lo = 0, hi = nk - 1;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (cs[mid + 1].x > target) hi = mid; else lo = mid + 1; }
        const uint4 t = cs[lo];
        lut[b] = make_uint2(t.x | ((t.z & 0xfffu) << 20), t.y | ((t.z >> 12) << 21)); kidx[b] = (uint16_t)lo;
    }
    __syncwarp();
    if (nk == 1 && !early) {          // a single symbol with probability one: the coder state never moves and no byte is read
        const uint32_t sy = cs[0].z; const int32_t val = mode == 1 ? ((sy & 1) ? -(int32_t)(sy >> 1) - 1 : (int32_t)(sy >> 1)) : (int32_t)sy;
        if (mode == 0) for (uint32_t i = lane; i < count; i += 32) ((uint8_t *)out)[i] = (uint8_t)sy;
        else for (uint32_t i = lane; i < count; i += 32) ((int32_t *)out)[i] = val;
        return;
    }
    if (lane != 0 || count == 0) return;
    // ---- the run (lane 0).  Bytes are consumed back to front through a register window of aligned 32-bit words,
    // the next word always prefetched, so renormalisation never waits on memory.
    const uint8_t *data = file + s.data_off; const uint32_t nbytes = s.data_len;
    if (nbytes == 0) { if (!early) frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }
    const uint32_t prec = 1u << pb, lbase = prec * 4u, shift = pb - lb;
    const unsigned x = data[nbytes - 1] >> 6, k = x + 1;
    if (nbytes < k) { if (!early) frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }
    uint32_t st = 0;
    for (unsigned i = 0; i < k; i++) st |= (uint32_t)data[nbytes - k + i] << (8 * i);
    st &= (1u << (8 * k - 2)) - 1u;
    st += lbase;
    const uint8_t *p = data + (nbytes - k);                   // next byte to consume is p[-1]
    const uint32_t *wp = (const uint32_t *)((uintptr_t)(p - 1) & ~(uintptr_t)3);     // aligned word holding it (the blob is padded on both sides)
    uint32_t w = p > data ? wp[0] : 0u, wn = p > data ? wp[-1] : 0u;
    uint32_t bi = (uint32_t)((uintptr_t)(p - 1) & 3);         // byte index inside w
    int left = (int)(nbytes - k);                             // bytes not yet consumed
    RansRun r{cs, lut, kidx, wp, w, wn, bi, left, st, pb, shift};
    const bool wide = A > 4096u;
    if (early) {
        uint32_t got;
        if (!wide) got = mode == 1 ? rans_walk<1, false, true>(r, count, out) : rans_walk<2, false, true>(r, count, out);
        else got = mode == 1 ? rans_walk<1, true, true>(r, count, out) : rans_walk<2, true, true>(r, count, out);
        counts[jb.frame].rans_early[jb.what - 16] = got;
    } else if (!wide) { if (mode == 0) rans_walk<0, false, false>(r, count, out); else if (mode == 1) rans_walk<1, false, false>(r, count, out); else rans_walk<2, false, false>(r, count, out); }
    else { if (mode == 0) rans_walk<0, true, false>(r, count, out); else if (mode == 1) rans_walk<1, true, false>(r, count, out); else rans_walk<2, true, false>(r, count, out); }
}

// TAGGED symbol scheme, second half (draco::DecodeTaggedSymbols): the tags k_rans decoded (one bit length per entry) -> exclusive prefix
// sum -> every entry's bit offset -> its components, read as LSB-first bit fields from the bytes behind the tag run and converted
// like RAW symbols (zig-zag unless the transform yields positive corrections).  grid = (frames, attributes), 256 threads.
__global__ void __launch_bounds__(256) k_tagged_values(const DracoFrame *frames, DracoCounts *counts, const uint8_t *blob, uint8_t *S) {
    __shared__ unsigned long long wsum[8], carry_s;
    const uint32_t fi = blockIdx.x, j = blockIdx.y;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if ((int)j >= f.nattr || f.o_tags[j] == UVOL_NONE || f.o_corr[j] == UVOL_NONE) return;
    const DracoAttr &a = f.attr[j];
    const uint32_t n = counts[fi].expected[a.table + 1], nc = (uint32_t)a.vnc;
    if (n > f.table_cap[a.table + 1]) { if (threadIdx.x == 0) frame_fail(counts, fi, UVOL_ERR_FRAME_CAPACITY); return; }
    const uint32_t *tags = (const uint32_t *)(S + f.o_tags[j]); int32_t *out = (int32_t *)(S + f.o_corr[j]);
    const uint8_t *bits = blob + f.file_off + a.tag_bits_off; const unsigned long long total = 8ull * a.tag_bits_len;
    const bool positive = a.pred != -2 && (a.xform == 2 || a.xform == 3);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    bool bad = false;
    for (uint32_t base = 0; base < n; base += 256) {
        const uint32_t e = base + tid;
        uint32_t t = e < n ? tags[e] : 0u;
        if (t > 32u) { bad = true; t = 0; }
        unsigned long long inc = (unsigned long long)t * nc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += x; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        unsigned long long before = carry_s;
        for (int k = 0; k < w; k++) before += wsum[k];
        const unsigned long long off = before + inc - (unsigned long long)t * nc;
        if (e < n) {
            if (off + (unsigned long long)t * nc > total) bad = true;
            else for (uint32_t c = 0; c < nc; c++) {
                const unsigned long long p = off + (unsigned long long)c * t; const uint8_t *b = bits + (p >> 3);
                unsigned long long x = 0;                                     // (the blob is padded: five bytes are always readable)
#pragma unroll
                for (int k = 0; k < 5; k++) x |= (unsigned long long)b[k] << (8 * k);
                const uint32_t v = t == 0 ? 0u : (uint32_t)(x >> (p & 7)) & (t == 32 ? 0xffffffffu : ((1u << t) - 1u));
                out[e * nc + c] = positive ? (int32_t)v : ((v & 1u) ? -(int32_t)(v >> 1) - 1 : (int32_t)(v >> 1));
            }
        }
        __syncthreads();
        if (tid == 255) carry_s = before + inc;
        __syncthreads();
    }
    if (bad) frame_fail(counts, fi, UVOL_ERR_CORRUPT);
}

// rABS bit runs, one run per lane.  what: 0..3 = seam bits of attribute data i (upper bound 3F/2+1 bits); 16+j = attribute j aux
// bits (TEX_COORDS orientations incl. the toggle decoding, or GEOMETRIC_NORMAL flip bits).  One byte per bit out, eight at a time.
__global__ void __launch_bounds__(32) k_rabs_lanes(const DracoFrame *frames, DracoCounts *counts, const uint8_t *blob,
                                                   uint8_t *scratch, const Job *jobs, int njobs) {
    const unsigned FULL = 0xffffffffu;
    const int ji = blockIdx.x * 32 + (int)threadIdx.x;
    bool live = ji < njobs;
    Job jb = {0u, 0};
    if (live) { jb = jobs[ji]; live = !frame_dead(frames, counts, jb.frame); }
    RabsLane r; rabs_lane_idle(r, blob); uint8_t *o = nullptr; uint32_t n = 0; bool toggle = false;
    if (live) {
        const DracoFrame &f = frames[jb.frame];
        const uint8_t *file = blob + f.file_off;
        bool ok;
        if (jb.what < 16) { ok = rabs_lane_init(r, file, f.seams[jb.what]); o = scratch + f.o_seambits[jb.what]; n = 3 * f.nf / 2 + 1; }
        else {
            const DracoAttr &a = f.attr[jb.what - 16];
            ok = rabs_lane_init(r, file, a.aux_bits); o = scratch + f.o_auxbits[jb.what - 16];
            n = counts[jb.frame].expected[a.table + 1];
            if (n > f.table_cap[a.table + 1]) { frame_fail(counts, jb.frame, UVOL_ERR_FRAME_CAPACITY); live = false; }
            else if (a.pred == 5) { if ((uint32_t)a.num_orient > n) ok = false; n = (uint32_t)a.num_orient; toggle = true; }
        }
        if (live && !ok) { frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); live = false; }
    }
    if (!live) n = 0;
    uint32_t last = 1;
    for (uint32_t k = 0; __any_sync(FULL, k < n); k += 8) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const bool on = k + b < n;
            uint32_t bit = rabs_lane_bit(r, on);
            last ^= (toggle && on && !bit) ? 1u : 0u;
            bit = on ? (toggle ? last : bit) : 0u;
            if (b < 4) lo |= bit << (8 * b); else hi |= bit << (8 * (b - 4));
        }
        if (k < n) *(uint2 *)(o + k) = make_uint2(lo, hi);      // (the arrays end on 8 bytes of slack)
    }
}

// Edgebreaker connectivity: one warp per frame, lane 0 walks the symbol sequence.
__global__ void __launch_bounds__(32 * SERIAL_WARPS) k_edgebreaker(const DracoFrame *frames, DracoCounts *counts, const uint8_t *blob, const uint32_t *aux,
                                                    uint8_t *S, int nframes) {
    const uint32_t fi = blockIdx.x * SERIAL_WARPS + (threadIdx.x >> 5);
    if ((int)fi >= nframes || (threadIdx.x & 31) != 0) return;
    if (frames[fi].status) { counts[fi].status = frames[fi].status; return; }
    if (counts[fi].status || frames[fi].trav == 2) return;      // valence frames: k_edgebreaker_valence2
    const DracoFrame &f = frames[fi];
    EbMem m; m.opp = (int *)(S + f.o_opp); m.c2v = (int *)(S + f.o_c2v); m.lmc = (int *)(S + f.o_lmc); m.val = (int *)(S + f.o_val);
    m.hole = S + f.o_hole; m.stack = (int *)(S + f.o_stack); m.skey = m.stack + f.nsym + 8; m.sval = m.skey + f.nts + 1; m.invalid = (int *)(S + f.o_invalid);
    for (int i = 0; i < 6; i++) m.ctxsym[i] = S + f.o_ctxsym[i];
    uint32_t slots = 0, valid = 0;
    const int rc = eb_decode_frame(f, blob + f.file_off, aux, m, &slots, &valid);
    counts[fi].num_vertex_slots = slots; counts[fi].expected[0] = valid;
    if (rc) frame_fail(counts, fi, rc);
}

// ---------------------------------------------------------------------------------------------
// Valence-mode edgebreaker (the generic k_edgebreaker above remains the path for the standard traversal).  The walk is a
// serial state machine (the next symbol's context depends on the mesh built so far), so the only lever is the length of the
// dependent chain per symbol.  The whole warp runs the state machine redundantly
// (uniform control flow costs nothing extra) so that lanes can specialise where it saves
// instructions on the serial chain:
//   * lanes 0..5 each own one context's symbol stream (eight symbols packed in a 64-bit register, the
//     next aligned word already prefetched); the context -> symbol step is one shuffle;
//   * only the two gate vertices' records live in registers (the tip never changes while it is the
//     tip, so it is written back the moment a vertex becomes the tip: one record store per symbol);
//   * the record keeps next(left-most corner) instead of the corner itself, so no modulo appears on
//     the chain; the gate corner is always corner 0 of the previous face and is not state at all;
//   * faces are staged as 16-byte records in shared memory and written to opp / c2v 32 at a time,
//     one face per lane (own links, then -- after a warp barrier -- the links into older faces);
//   * consistency checks accumulate in a sticky flag (all indices stay in range whatever the stream
//     says) that is examined when the walk ends.
// S symbols (component merges) flush the staged faces and use the memory path, executed uniformly.
// (compute-sanitizer racecheck reports the redundant execution as hazards on the ring / stage words: every lane writes the SAME
// value to the same word and reads back what it -- or a lane in the same state -- wrote; memcheck is clean.)
#define EB2_RING 1024
#define EB2_STAGE 32
__global__ void __launch_bounds__(32 * SERIAL_WARPS) k_edgebreaker_valence2(const DracoFrame *frames, DracoCounts *counts, const uint8_t *blob, const uint32_t *aux,
                                                             uint8_t *S, int nframes) {
    extern __shared__ uint4 eb_smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 *ring = eb_smem + (size_t)wib * (EB2_RING + 128 + EB2_STAGE);
    int *stk = (int *)(ring + EB2_RING);
    uint4 *stage = ring + EB2_RING + 128;
    const uint32_t fi = blockIdx.x * SERIAL_WARPS + wib;
    if ((int)fi >= nframes) return;
    if (frames[fi].status) { if (lane == 0) counts[fi].status = frames[fi].status; return; }
    if (counts[fi].status || frames[fi].trav != 2) return;
    const DracoFrame &f = frames[fi];
    int *opp = (int *)(S + f.o_opp), *c2v = (int *)(S + f.o_c2v), *lmc = (int *)(S + f.o_lmc), *gstk = (int *)(S + f.o_stack);
    uint8_t *hole = S + f.o_hole; uint4 *vrec = (uint4 *)(S + f.o_val);     // o_val is sized 16 B per vertex slot (see draco_plan.h)
    int *skey = gstk + f.nsym + 8, *sval = skey + f.nts + 1, *invalid = (int *)(S + f.o_invalid);
    const int F = (int)f.nf, maxv = (int)(f.nv_enc + f.nsplit), nsym = (int)f.nsym;
    int nverts = 0, ninv = 0, sp = 0, status = 0;
    if (nsym > F || nsym < 1) status = UVOL_ERR_CORRUPT;

    // ---- per-lane symbol stream (lanes 0..5); lanes >= 6 answer E forever (used for the very first symbol).
    // w holds up to eight symbols, pos is the byte index of the next one (counting down), cur the symbol itself.
    uint32_t wlo = 0x04040404u, whi = 0x04040404u; unsigned long long nxt = 0; int pos = 0x40000007, blk = 0; const unsigned long long *cs8 = nullptr;
    if (lane < 6) {
        const int kk = (int)f.ctx[lane].count; cs8 = (const unsigned long long *)(S + f.o_ctxsym[lane]);
        if (kk <= 0) { wlo = whi = ~0u; }
        else {
            blk = (kk - 1) >> 3; pos = kk - 8 * blk - 1;
            const unsigned long long w0 = cs8[blk]; wlo = (uint32_t)w0; whi = (uint32_t)(w0 >> 32);
            if (blk > 0) nxt = cs8[blk - 1];
        }
    }
    int cur = (int)(__byte_perm(wlo, whi, (uint32_t)pos) & 255u);
    const uint32_t *ts = aux + f.ts_off; int ts_top = (int)f.nts, nsa = 0;
    int next_ts_sid = ts_top > 0 ? nsym - 1 - (int)ts[3 * (ts_top - 1)] : -1;
    // record: x = next(left-most corner) (or DINV), y = vertex before it on the boundary, z = 2 * valence + on-hole flag
    int na = 0, pa = 0; int rn_c = 0, rn_p = 0, rn_v = 0, rp_c = 0, rp_p = 0, rp_v = 0;
    int ctx = 6, lo = -EB2_RING, fbase = 0, pend = 0;
#define VLOAD2(v, C_, P_, V_) do { uint4 u_; if ((v) >= lo) u_ = ring[(v) & (EB2_RING - 1)]; else u_ = vrec[(v)]; C_ = (int)u_.x; P_ = (int)u_.y; V_ = (int)u_.z; } while (0)
#define VSTORE2(v, C_, P_, V_) do { const uint4 u_ = make_uint4((uint32_t)(C_), (uint32_t)(P_), (uint32_t)(V_), 0u); if ((v) >= lo) ring[(v) & (EB2_RING - 1)] = u_; vrec[(v)] = u_; } while (0)
#define EB2_FLUSH() do { \
        int s_ = 4, x_ = DINV, c_ = 0; \
        if (lane < pend) { const uint4 r_ = stage[lane]; c_ = 3 * (fbase + lane); s_ = (int)(r_.w >> 28); x_ = (int)(r_.w & 0x0fffffffu); \
            c2v[c_] = (int)r_.x; c2v[c_ + 1] = (int)r_.y; c2v[c_ + 2] = (int)r_.z; \
            opp[c_] = DINV; opp[c_ + 1] = (s_ == 0 || s_ == 2) ? c_ - 3 : DINV; opp[c_ + 2] = s_ == 0 ? x_ : (s_ == 3 ? c_ - 3 : DINV); } \
        __syncwarp(); \
        if (lane < pend && s_ < 4) { opp[c_ - 3] = s_ == 3 ? c_ + 2 : c_ + 1; if (s_ == 0 && x_ < c_) opp[x_] = c_ + 2; } \
        __syncwarp(); \
        fbase += pend; pend = 0; } while (0)
    int sid = 0;
    for (; sid < nsym && !status; sid++) {
        const int s = __shfl_sync(0xffffffffu, cur, ctx);
        {
            const bool mine = lane == ctx;
            if (mine && (pos & 7) == 0) {               // that was the last symbol of this word (one lane, every eighth symbol)
                if (blk == 0) { wlo = whi = ~0u; pos = 0x40000008; }
                else { wlo = (uint32_t)nxt; whi = (uint32_t)(nxt >> 32); pos = 8; --blk; if (blk > 0) nxt = cs8[blk - 1]; }
            }
            if (mine) pos--;
            cur = (int)(__byte_perm(wlo, whi, (uint32_t)pos) & 255u);
        }
        const int c0 = 3 * sid;
        int s_rl = s, s_rare = s;                       // opaque copies: keep the dispatch a compare chain ordered by frequency, not a jump table
        asm volatile("" : "+r"(s_rl)); asm volatile("" : "+r"(s_rare));
        if (s == 0) {              // C: close the fan at the vertex next to the gate
            const int vbn = rn_p, b = rn_c;
            if ((b < 0) | (b == c0 - 3) | (vbn == na) | (vbn == pa)) { status = UVOL_ERR_CORRUPT; break; }
            int rb_c, rb_p, rb_v; VLOAD2(vbn, rb_c, rb_p, rb_v);
            stage[pend] = make_uint4((uint32_t)na, (uint32_t)vbn, (uint32_t)pa, (uint32_t)(b & 0x0fffffff));
            rp_c = c0; rp_p = vbn; rp_v += 2;
            rn_v &= ~1; VSTORE2(na, rn_c, rn_p, rn_v);
            na = vbn; rn_c = rb_c; rn_p = rb_p; rn_v = rb_v + 2;
        } else if (s_rl == 3) {                         // R: new vertex opposite the gate, continue to the left
            if (nverts >= maxv) { status = UVOL_ERR_CORRUPT; break; }
            const int nvx = nverts++; lo = nverts - EB2_RING;
            stage[pend] = make_uint4((uint32_t)pa, (uint32_t)na, (uint32_t)nvx, 3u << 28);
            rp_c = c0 + 1; rp_p = nvx; rp_v += 2; VSTORE2(pa, rp_c, rp_p, rp_v);
            rn_v += 2;
            pa = nvx; rp_c = c0; rp_p = na; rp_v = 5;
        } else if (s_rl == 2) {                         // L
            if (nverts >= maxv) { status = UVOL_ERR_CORRUPT; break; }
            const int nvx = nverts++; lo = nverts - EB2_RING;
            stage[pend] = make_uint4((uint32_t)na, (uint32_t)nvx, (uint32_t)pa, 2u << 28);
            rp_c = c0; rp_p = nvx; rp_v += 2;
            rn_v += 2; VSTORE2(na, rn_c, rn_p, rn_v);
            rn_c = c0 + 2; rn_p = na; rn_v = 5; na = nvx;
        } else if (s_rare == 4) {                       // E: isolated triangle, the old gate goes on the stack
            if (nverts + 3 > maxv) { status = UVOL_ERR_CORRUPT; break; }
            const int v0 = nverts, v1 = nverts + 1, v2 = nverts + 2; nverts += 3; lo = nverts - EB2_RING;
            if (sid > 0) { VSTORE2(na, rn_c, rn_p, rn_v); VSTORE2(pa, rp_c, rp_p, rp_v); }
            stage[pend] = make_uint4((uint32_t)v0, (uint32_t)v1, (uint32_t)v2, 4u << 28);
            VSTORE2(v0, c0 + 1, v2, 5);
            na = v1; rn_c = c0 + 2; rn_p = v0; rn_v = 5;
            pa = v2; rp_c = c0; rp_p = v1; rp_v = 5;
            if (sp > 0) { if (sp <= 512) stk[sp - 1] = c0 - 3; else gstk[sp - 1] = c0 - 3; }
            sp++;
        } else if (s_rare == 1) {                       // S: merge the two topmost components (memory path on lane 0, state broadcast)
            VSTORE2(na, rn_c, rn_p, rn_v); VSTORE2(pa, rp_c, rp_p, rp_v);
            EB2_FLUSH();
            fbase = sid + 1;                            // this face is written directly below
            if (lane == 0) do {
                const int b = c0 - 3; sp--;
                int a2 = -1;
                for (int k = 0; k < nsa; k++) if (skey[k] == sid) { a2 = sval[k]; sp++; break; }
                if (a2 < 0) { if (sp == 0) { status = UVOL_ERR_CORRUPT; break; } a2 = sp <= 512 ? stk[sp - 1] : gstk[sp - 1]; }
                if (a2 == b || opp[a2] >= 0 || opp[b] >= 0) { status = UVOL_ERR_CORRUPT; break; }
                const int vp = c2v[cprev(a2)], vnx = c2v[cnext(a2)], vbp = pa, vn = na;
                opp[c0] = DINV; opp[c0 + 2] = a2; opp[a2] = c0 + 2; opp[c0 + 1] = b; opp[b] = c0 + 1;
                c2v[c0] = vp; c2v[c0 + 1] = vnx; c2v[c0 + 2] = vbp;
                int t_c, t_p, t_v;
                VLOAD2(vbp, t_c, t_p, t_v); VSTORE2(vbp, c0, vnx, t_v);
                int p_c, p_p, p_v, n_c, n_p, n_v;
                VLOAD2(vp, p_c, p_p, p_v); VLOAD2(vn, n_c, n_p, n_v);
                p_v = ((p_v >> 1) + (n_v >> 1)) * 2 + (p_v & 1); VSTORE2(vp, n_c, n_p, p_v);
                int cn = cnext(b); const int first = cn; int guard = 0;
                while (cn >= 0) {
                    c2v[cn] = vp;
                    const int cx = cnext(cn), x = c2v[cx];
                    if (x != vn) { int x_c, x_p, x_v; VLOAD2(x, x_c, x_p, x_v); if (x_c == cnext(cx)) VSTORE2(x, x_c, vp, x_v); }
                    cn = b_swl(opp, cn);
                    if (cn == first || ++guard > 3 * F) { status = UVOL_ERR_CORRUPT; break; }
                }
                if (status) break;
                VSTORE2(vn, DINV, n_p, n_v);
                invalid[ninv] = vn;
                na = vnx; pa = vbp;
                VLOAD2(na, rn_c, rn_p, rn_v); VLOAD2(pa, rp_c, rp_p, rp_v);
                rn_v += 2; rp_v += 2;                   // the new tip (vp) is already in memory
            } while (0);
            __syncwarp();
            ninv++;
            status = __shfl_sync(0xffffffffu, status, 0); sp = __shfl_sync(0xffffffffu, sp, 0);
            na = __shfl_sync(0xffffffffu, na, 0); pa = __shfl_sync(0xffffffffu, pa, 0);
            rn_c = __shfl_sync(0xffffffffu, rn_c, 0); rn_p = __shfl_sync(0xffffffffu, rn_p, 0); rn_v = __shfl_sync(0xffffffffu, rn_v, 0);
            rp_c = __shfl_sync(0xffffffffu, rp_c, 0); rp_p = __shfl_sync(0xffffffffu, rp_p, 0); rp_v = __shfl_sync(0xffffffffu, rp_v, 0);
            if (status) break;
            pend = -1;                                  // ++ below makes it 0: nothing staged for this face
        } else { status = UVOL_ERR_CORRUPT; break; }
        if (++pend == EB2_STAGE) EB2_FLUSH();
        { int v = rn_v >> 1; v = v < 2 ? 2 : (v > 7 ? 7 : v); ctx = v - 2; }
        if (sid == next_ts_sid) {                       // topology split events registered on this symbol (A.3)
            if (s >= 2) {
                while (ts_top > 0 && ts[3 * (ts_top - 1)] == (uint32_t)(nsym - sid - 1)) {
                    --ts_top;
                    if (lane == 0) { skey[nsa] = nsym - (int)ts[3 * ts_top + 1] - 1; sval[nsa] = ts[3 * ts_top + 2] == 1 ? c0 + 1 : c0 + 2; }
                    nsa++;
                }
                __syncwarp();
            }
            next_ts_sid = ts_top > 0 ? nsym - 1 - (int)ts[3 * (ts_top - 1)] : -1;
        }
    }
    if (!status) {
        VSTORE2(na, rn_c, rn_p, rn_v); VSTORE2(pa, rp_c, rp_p, rp_v);
        EB2_FLUSH();
        if (sp > 0) { if (sp <= 512) stk[sp - 1] = 3 * (nsym - 1); else gstk[sp - 1] = 3 * (nsym - 1); }
    }
    __syncwarp();
#undef VLOAD2
#undef VSTORE2
    int numf = nsym;
    // records -> the plain arrays the later kernels read (all lanes)
    if (!status) for (int v = lane; v < nverts; v += 32) { const uint4 u = vrec[v]; const int c = (int)u.x; lmc[v] = c < 0 ? DINV : cprev(c); hole[v] = (uint8_t)(u.z & 1u); }
    __syncwarp();
    if (lane != 0) return;
    if (!status) {
        // start faces (one rABS bit per remaining stack entry), then fold isolated vertices away
        if (sp > 0) {
            Rabs sf;
            if (!rabs_init(sf, blob + f.file_off, f.start_faces)) status = UVOL_ERR_CORRUPT;
            while (!status && sp > 0) {
                const int corner = sp <= 512 ? stk[sp - 1] : gstk[sp - 1]; sp--;
                if (rabs_bit(sf)) {
                    const int a = corner, vn = c2v[cnext(a)];
                    if (lmc[vn] < 0) { status = UVOL_ERR_CORRUPT; break; }
                    const int cb = cnext(lmc[vn]), vx = c2v[cnext(cb)];
                    if (lmc[vx] < 0) { status = UVOL_ERR_CORRUPT; break; }
                    const int cc = cnext(lmc[vx]);
                    if (a == cb || cb == cc || a == cc || opp[a] >= 0 || opp[cb] >= 0 || opp[cc] >= 0 || numf >= F) { status = UVOL_ERR_CORRUPT; break; }
                    const int vp = c2v[cnext(cc)], nc = 3 * numf++;
                    opp[nc] = a; opp[a] = nc; opp[nc + 1] = cb; opp[cb] = nc + 1; opp[nc + 2] = cc; opp[cc] = nc + 2;
                    c2v[nc] = vx; c2v[nc + 1] = vp; c2v[nc + 2] = vn;
                    hole[vx] = 0; hole[vp] = 0; hole[vn] = 0;
                }
            }
        }
        if (!status && numf != F) status = UVOL_ERR_CORRUPT;
        if (!status) {
            int num_vertices = nverts;
            for (int k = 0; k < ninv && !status; k++) {
                const int iv = invalid[k];
                int src = num_vertices - 1;
                while (src >= 0 && lmc[src] == DINV) src = --num_vertices - 1;
                if (src < iv) continue;
                const int cs = lmc[src]; int c = cs, left = 1, guard = 0;
                while (c >= 0) {
                    int nx;
                    if (left) { nx = b_swl(opp, c); if (nx < 0) { nx = b_swr(opp, cs); left = 0; } else if (nx == cs) nx = DINV; }
                    else nx = b_swr(opp, c);
                    if (c2v[c] != src || ++guard > 3 * F) { status = UVOL_ERR_CORRUPT; break; }
                    c2v[c] = iv; c = nx;
                }
                lmc[iv] = lmc[src]; lmc[src] = DINV;
                hole[iv] = hole[src]; hole[src] = 0;
                num_vertices--;
            }
        }
    }
    counts[fi].num_vertex_slots = (uint32_t)nverts; counts[fi].expected[0] = (uint32_t)(nverts - ninv);
    if (status) frame_fail(counts, fi, status);
}

// Attribute seams.  The k-th bit of each seam stream belongs to the k-th corner (in corner order) whose opposite face is not
// older than its own, so a corner needs the number of such corners before it: k_seam_count totals them per chunk of
// SEAM_CHUNK corners, k_seams adds up the chunks before its own (at most a few hundred) and runs a ballot-based block
// scan inside the chunk.  grid = (ceil(3 * maxF / SEAM_CHUNK), frames), 256 threads.
#define SEAM_CHUNK 8192
__global__ void __launch_bounds__(256) k_seam_count(const DracoFrame *frames, const DracoCounts *counts, uint8_t *S) {
    __shared__ int warp_tot[8];
    const uint32_t fi = blockIdx.y;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    const int C = 3 * (int)f.nf, base = (int)blockIdx.x * SEAM_CHUNK, tid = threadIdx.x;
    if (f.nad == 0 || base >= C) return;
    const int *opp = (const int *)(S + f.o_opp);
    int mine = 0;
    for (int c = base + tid; c < min(C, base + SEAM_CHUNK); c += 256) { const int o = opp[c]; mine += (o >= 0 && o / 3 >= c / 3); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
    if ((tid & 31) == 0) warp_tot[tid >> 5] = mine;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int k = 0; k < 8; k++) t += warp_tot[k]; ((int *)(S + f.o_seamcnt))[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(256) k_seams(const DracoFrame *frames, DracoCounts *counts, uint8_t *S, uint8_t *Z, int nframes) {
    __shared__ int warp_tot[8]; __shared__ int carry;
    const uint32_t fi = blockIdx.y;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    const int C = 3 * (int)f.nf, nad = (int)f.nad, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int chunk0 = (int)blockIdx.x * SEAM_CHUNK;
    if (nad == 0 || chunk0 >= C) return;
    const int *opp = (const int *)(S + f.o_opp), *c2v = (const int *)(S + f.o_c2v);
    {   // flagged corners in the chunks before this one
        const int *cnt = (const int *)(S + f.o_seamcnt); int mine = 0;
        for (int k = tid; k < (int)blockIdx.x; k += 256) mine += cnt[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
        if (lane == 0) warp_tot[w] = mine;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int k = 0; k < 8; k++) t += warp_tot[k]; carry = t; }
        __syncthreads();
    }
    for (int base = chunk0; base < min(C, chunk0 + SEAM_CHUNK); base += 256) {
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
    if (tid == 0) { if (which < 4) { counts[fi].attr_vertices[which] = (uint32_t)carry_s; counts[fi].expected[1 + which] = (uint32_t)carry_s; } else counts[fi].num_points = (uint32_t)carry_s; }
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

// Device planner: lays out the count-sized arrays of every frame (point -> corner map, entry maps, per-point output arrays) from
// the counts the connectivity kernels have just produced, with the same function the host runs on the final counts
// (draco_plan2_frame): per-frame sizes, block-wide exclusive scan over the frames, offsets written into the device descriptors.
// A frame whose attribute tables outgrew their optimistic capacity is failed with UVOL_ERR_FRAME_CAPACITY, a batch that outgrew the
// reserved arenas fails every frame with UVOL_ERR_BATCH_CAPACITY (the launcher re-plans and runs the batch again).  One block of 256
// threads: small enough to start at once next to the resident entropy / connectivity blocks of the side streams (a 1024-thread
// block waited 14 ms for an SM to drain).
#define PLAN_T 256
__global__ void __launch_bounds__(PLAN_T) k_plan2(DracoFrame *frames, DracoCounts *counts, DracoBatchPlan *bp, int n, uint64_t out_index_bytes,
                                                uint64_t cap_s2, uint64_t cap_z2, uint64_t cap_out) {
    __shared__ unsigned long long wsum[6][PLAN_T / 32]; __shared__ unsigned long long carry[6], total[4]; __shared__ int over;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // pass 0 totals the per-slot output sizes (the region bases need them), pass 1 assigns
    for (int pass = 0; pass < 2; pass++) {
        if (tid == 0) {
            carry[0] = carry[1] = 0;
            if (pass == 0) { for (int k = 0; k < 4; k++) carry[2 + k] = 0; }
            else { uint64_t base[4], need; const uint64_t t4[4] = {total[0], total[1], total[2], total[3]}; draco_plan2_regions(out_index_bytes, t4, base, &need);
                   for (int k = 0; k < 4; k++) { carry[2 + k] = base[k]; bp->slot_base[k] = base[k]; bp->slot_bytes[k] = total[k]; } bp->out_need = need; }
        }
        __syncthreads();
        for (int base = 0; base < n; base += PLAN_T) {
            const int i = base + tid;
            Plan2Cursor sz{0, 0, {0, 0, 0, 0}};
            if (i < n && !frame_dead(frames, counts, i)) {
                if (pass == 0) for (uint32_t t = 1; t <= frames[i].nad; t++) if (counts[i].attr_vertices[t - 1] > frames[i].table_cap[t]) counts[i].status = UVOL_ERR_FRAME_CAPACITY;
                draco_plan2_frame(frames[i], counts[i], sz, false);
            }
            unsigned long long x[6] = {sz.s, sz.z, sz.o[0], sz.o[1], sz.o[2], sz.o[3]}, inc[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                unsigned long long v = x[k];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
                inc[k] = v;
                if (lane == 31) wsum[k][w] = v;
            }
            __syncthreads();
            unsigned long long pre[6];
#pragma unroll
            for (int k = 0; k < 6; k++) { pre[k] = carry[k]; for (int j = 0; j < w; j++) pre[k] += wsum[k][j]; pre[k] += inc[k] - x[k]; }
            if (pass == 1 && i < n && !frame_dead(frames, counts, i)) {
                Plan2Cursor cur{pre[0], pre[1], {pre[2], pre[3], pre[4], pre[5]}};
                draco_plan2_frame(frames[i], counts[i], cur, true);
            }
            __syncthreads();
            if (tid == PLAN_T - 1) for (int k = 0; k < 6; k++) carry[k] = pre[k] + x[k];
            __syncthreads();
        }
        if (tid == 0 && pass == 0) for (int k = 0; k < 4; k++) total[k] = carry[2 + k];
        __syncthreads();
    }
    if (tid == 0) {
        bp->s2_need = carry[0]; bp->z2_need = carry[1];
        over = carry[0] > cap_s2 || carry[1] > cap_z2 || bp->out_need > cap_out;
        bp->overflow = (uint32_t)over;
    }
    __syncthreads();
    if (over) for (int i = tid; i < n; i += PLAN_T) if (!frames[i].status && !counts[i].status) counts[i].status = UVOL_ERR_BATCH_CAPACITY;
}

// Traversal records of one corner table (element-parallel, one thread per face): the three vertex ids, the three opposite corners
// (cut at seams), boundary flags, the corners through which a walk along the face-id order enters the face (from f-1 / from f+1) and
// -- for those two entry corners -- the distance to the nearest earlier face of the run whose entry corner has the same tip vertex
// (so the traversal knows, without any exchange between lanes, that a tip was already reached inside the same 32-face step).
// 32 bytes per face; everything the serial traversal needs about a face is one load.   grid = (ceil(maxF/256), traversal jobs)
// which: 0 = base tables only, 1 = attribute tables only, 2 = all
__global__ void __launch_bounds__(256) k_face_records(const DracoFrame *frames, const DracoCounts *counts, uint8_t *S, const uint8_t *Z, const Job *jobs, int which) {
    __shared__ uint32_t tipU[256 + 31], tipD[256 + 31];
    const Job jb = jobs[blockIdx.y];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame]; const int t = jb.what;
    if (f.o_frec[t] == UVOL_NONE || (which == 0 && t != 0) || (which == 1 && t == 0)) return;
    const int F = (int)f.nf, base = blockIdx.x * 256, tid = threadIdx.x;
    if (base >= F) return;
    const TableView tv = make_view(f, t, S, Z);
    const int *lmc = (const int *)(S + f.o_lmc);
    const int fi = base + tid;
    FaceRec r; r.meta = 0xfu; uint32_t mu = 0xffffffffu, md = 0xffffffffu;
    if (fi < F) {
        face_record(fi, tv, lmc, r);
        const uint32_t ku = FREC_UPK(r.meta), kd = FREC_DNK(r.meta);
        if (ku < 3) mu = (uint32_t)(ku == 0 ? r.v[0] : (ku == 1 ? r.v[1] : r.v[2]));
        if (kd < 3) md = (uint32_t)(kd == 0 ? r.v[0] : (kd == 1 ? r.v[1] : r.v[2]));
    }
    tipU[tid + 31] = mu; tipD[tid] = md;
    if (tid < 31) { tipU[tid] = face_entry_tip(base - 31 + tid, F, 1, tv); tipD[256 + tid] = face_entry_tip(base + 256 + tid, F, -1, tv); }
    __syncthreads();
    if (fi >= F) return;
    uint32_t du = 0, dd = 0;
    if (mu != 0xffffffffu) { for (int k = 1; k < 32; k++) if (tipU[tid + 31 - k] == mu) { du = (uint32_t)k; break; } }
    if (md != 0xffffffffu) { for (int k = 1; k < 32; k++) if (tipD[tid + k] == md) { dd = (uint32_t)k; break; } }
    r.meta |= (du << 4) | (dd << 9);
    uint4 *dst = (uint4 *)(S + f.o_frec[t]) + 2 * (size_t)fi;
    dst[0] = make_uint4((uint32_t)r.v[0], (uint32_t)r.v[1], (uint32_t)r.v[2], (uint32_t)r.o[0]);
    dst[1] = make_uint4((uint32_t)r.o[1], (uint32_t)r.o[2], r.meta, 0u);
}

// Depth-first traversal (A.3): one warp per (frame, table), up to 32 faces per step.
// Face ids follow the edgebreaker strip order and the traversal mostly walks along the same strips, so lane i SPECULATES which face
// the walk reaches i moves from now and through which corner:
//   * default: face f0 + i*dir through that face's static entry corner (dir = the direction of the last move; lane 0 stands on
//     the walk's actual corner);
//   * row pattern: when the last steps each ended after k faces with a jump of D faces (a strip crossed sideways: an attribute table
//     cut by a seam sends the walk across the rings of a UV sphere, two faces per ring), face f0 + (i / k)*D + (i % k)*dir; a lane
//     at the start of a row takes as entry corner the corner of its own record whose opposite lies in the previous lane's face.
// One 32-byte record load per lane (k_face_records), then every lane evaluates the exact step rule for its face against the
// visited maps plus the effects of the lanes before it (tips already reached inside the step: static duplicate distances along the
// face order, __match_any_sync on a row pattern; faces visited inside the step: index arithmetic on the pattern); the longest prefix whose transitions really lead to the next lane's corner is
// committed at once, and the first lane that deviates (pop, push, turn, direction change, jump) hands its exact outcome to the next
// step.  A wrong guess only shortens the step.  Output order is identical to the serial walk (tests/tools/draco_emu.cpp runs this
// very scheme lane by lane on the host against traverse_table).
#define TRAV_STACK 512
__device__ __forceinline__ unsigned face_of(int c) { return __umulhi((unsigned)c, 0xAAAAAAABu) >> 1; }
// GMAP = 0: visited-face / visited-vertex bitmaps in shared memory (F/8 + V/8 bytes per walk: fastest while all walks of the
// batch are co-resident).  GMAP = 1: a byte per face in global memory plus the vertex -> entry map itself as the visited-vertex
// test -- 2 KB of shared memory per walk, so large meshes (C3: 77 KB of bitmaps per walk) no longer cap the SM at two walks.
// Only this warp touches those bytes, so plain (L1-cached) loads / stores ordered by warp barriers suffice.
// Is the face at distance d from the step's first face one of the faces of lanes 0..lane?  Default guess: d * dir in [0, lane].
// Row pattern (k faces per row, rows D apart, |D| > 32 >= k so the decomposition is unique): d = q * D + r * dir with 0 <= r < k.
__device__ __forceinline__ bool trav_in_step(int d, int lane, int pk, int pD, int pdir) {
    if (pk == 32) { const int k = d * pdir; return k >= 0 && k <= lane; }
    const int e = d * pdir, PD = pD * pdir;                      // r = e - q * PD
    int q = e / PD; if (e - q * PD < 0) q += PD > 0 ? -1 : 1;   // floor-style: the remainder must be non-negative
    const int r = e - q * PD;
    return q >= 0 && r >= 0 && r < pk && q * pk + r <= lane;
}
template <int GMAP>
__global__ void __launch_bounds__(32) k_traverse(const DracoFrame *frames, DracoCounts *counts, uint8_t *S, uint8_t *Z, uint8_t *S2, uint8_t *Z2,
                                                 const Job *jobs, int njobs, int fwords_max, int vwords_max) {
    extern __shared__ uint32_t sm[];
    const int ji = blockIdx.x;
    if (ji >= njobs) return;
    constexpr bool G = GMAP != 0;
    const int bitwords = G ? 0 : fwords_max + vwords_max;
    const Job jb = jobs[ji];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame]; const int t = jb.what;
    if (f.o_frec[t] == UVOL_NONE || f.o_d2c[t] == UVOL_NONE) return;
    const int F = (int)f.nf, C = 3 * F, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t *fbits = sm, *vbits = sm + (G ? 0 : fwords_max); int *stk = (int *)(sm + bitwords);
    for (int i = lane; i < bitwords; i += 32) sm[i] = 0;
    const uint4 *grec = (const uint4 *)(S + f.o_frec[t]);
    int *d2c = (int *)(S2 + f.o_d2c[t]), *v2d1 = (int *)(Z2 + f.o_v2d[t]), *gst = (int *)(S + f.o_tstack[t]);
    uint8_t *fvis = Z + f.o_fvis[t];                         // GMAP only (zero-initialised with the arena)
    const int max_entries = (int)(t == 0 ? counts[jb.frame].num_vertex_slots : counts[jb.frame].attr_vertices[t - 1]);
    if (!G && (F > fwords_max * 32 || max_entries > vwords_max * 32)) { if (lane == 0) frame_fail(counts, jb.frame, UVOL_ERR_CORRUPT); return; }
    __syncwarp();
    int n = 0, sp = 0, c = -1, fscan = 0, status = 0, pdir = 1;
    int pk = 32, pD = 0, lastk = 0, lastD = 0;      // row pattern in use (pk = 32: none), row length / jump of the step before
    int prow = 0, pcol = lane;
#define FBIT(x) (G ? (uint32_t)fvis[(x)] : ((fbits[(x) >> 5] >> ((x) & 31)) & 1u))
#define VBIT(x) (G ? (uint32_t)(v2d1[(x)] != 0) : ((vbits[(x) >> 5] >> ((x) & 31)) & 1u))
    for (;;) {
        if (c < 0) {
            // ---- pick the next corner: stack top (lane 0), else the next unvisited face starts a component
            int done = 0, scan = 0;
            if (lane == 0) {
                for (;;) {
                    if (sp == 0) { scan = 1; break; }
                    c = sp <= TRAV_STACK ? stk[sp - 1] : gst[sp - 1];
                    if (c < 0 || c >= C || FBIT(face_of(c))) { sp--; c = -1; continue; }
                    break;
                }
            }
            scan = __shfl_sync(0xffffffffu, scan, 0);
            if (scan) {
                int nf = F;
                fscan = __shfl_sync(0xffffffffu, fscan, 0);
                if (G) {             // all lanes: 128 faces per round, four visited-bytes per lane
                    for (int base = fscan & ~127; base < F && nf == F; base += 128) {
                        const int i0 = base + lane * 4;
                        const uint32_t w = i0 < F ? *(const uint32_t *)(fvis + i0) : 0x01010101u;
                        int first = F;
#pragma unroll
                        for (int k = 3; k >= 0; k--) if (((w >> (8 * k)) & 255u) == 0 && i0 + k < F && i0 + k >= fscan) first = i0 + k;
                        const unsigned any = __ballot_sync(0xffffffffu, first < F);
                        if (any) nf = __shfl_sync(0xffffffffu, first, __ffs(any) - 1);
                    }
                } else if (lane == 0) {
                    int w = fscan >> 5; uint32_t m = ~fbits[w] & (0xffffffffu << (fscan & 31));
                    while (m == 0 && w + 1 < fwords_max) { w++; m = ~fbits[w]; }
                    nf = m ? w * 32 + __ffs(m) - 1 : F;
                }
                if (lane == 0) {
                    if (nf >= F) done = 1;
                    else {
                        fscan = nf; c = 3 * nf; stk[0] = c; sp = 1;
                        const uint4 A = grec[2 * (size_t)nf];
                        const int vn = (int)A.y, vp = (int)A.z;                                   // next / previous vertices first
                        if (!VBIT(vn) && n < max_entries) { if (!G) vbits[vn >> 5] |= 1u << (vn & 31); v2d1[vn] = ++n; d2c[n - 1] = c + 1; }
                        if (!VBIT(vp) && n < max_entries) { if (!G) vbits[vp >> 5] |= 1u << (vp & 31); v2d1[vp] = ++n; d2c[n - 1] = c + 2; }
                    }
                }
            }
            __syncwarp();
            done = __shfl_sync(0xffffffffu, done, 0);
            if (done) break;
            c = __shfl_sync(0xffffffffu, c, 0); n = __shfl_sync(0xffffffffu, n, 0); sp = __shfl_sync(0xffffffffu, sp, 0); fscan = __shfl_sync(0xffffffffu, fscan, 0);
        }
        if (c >= C) { status = UVOL_ERR_CORRUPT; break; }
        // ---- one speculative step over up to 32 faces
        const int f0 = (int)face_of(c), k0 = c - 3 * f0;
        const int fi = f0 + prow * pD + pcol * pdir;      // (prow, pcol: this lane's row / column under the pattern in use; 0, lane by default)
        const bool inr = fi >= 0 && fi < F;
        uint4 A = make_uint4(0, 0, 0, 0), B = make_uint4(0, 0, 0xfu, 0);
        if (inr) { A = grec[2 * (size_t)fi]; B = grec[2 * (size_t)fi + 1]; }
        int kf = lane == 0 ? k0 : -1;
        if (pk != 32) {   // the first lane of a later row: entered from the previous lane's face
            const int pf = __shfl_up_sync(0xffffffffu, fi, 1);
            if (lane > 0 && pcol == 0) {
                const int o0 = (int)A.w, o1 = (int)B.x, o2 = (int)B.y;
                kf = (o0 >= 0 && (int)face_of(o0) == pf) ? 0 : ((o1 >= 0 && (int)face_of(o1) == pf) ? 1 : ((o2 >= 0 && (int)face_of(o2) == pf) ? 2 : 3));
            }
        }
        const TravLane L = trav_lane((int)A.x, (int)A.y, (int)A.z, (int)A.w, (int)B.x, (int)B.y, B.z, fi, kf, pdir, inr && kf != 3);
        const int ci = L.ci, rc = L.rc, lc = L.lc; const unsigned v = L.v;
        const bool selfopen = ci >= 0 && !FBIT(fi);
        // vertex visited before my step: the map, or the tip of an earlier lane of this step -- along the face order that is lane 0's
        // actual tip or the static distance to an earlier face of the run with the same entry tip (k_face_records); on a row pattern
        // the lanes compare their tips (MATCH.ANY costs a round per distinct value, so it is kept off the default path)
        bool dup;
        if (pk == 32) { const unsigned v_first = __shfl_sync(0xffffffffu, v, 0); dup = lane > 0 && (v == v_first || (L.pd != 0 && (int)L.pd < lane)); }
        else dup = (__match_any_sync(0xffffffffu, ci >= 0 ? v : (0x80000000u | (unsigned)lane)) & lt) != 0;
        const bool vis = ci >= 0 && (VBIT(v) || dup);
        // neighbour faces visited before / during this step (faces of the lanes up to and including me)
        bool fr = true, fl = true;
        if (ci >= 0) {
            if (rc >= 0) { const int rf = (int)face_of(rc); fr = FBIT(rf) || trav_in_step(rf - f0, lane, pk, pD, pdir); }
            if (lc >= 0) { const int lf = (int)face_of(lc); fl = FBIT(lf) || trav_in_step(lf - f0, lane, pk, pD, pdir); }
        }
        int act, nx;
        trav_decide(vis, L.ob, fr, fl, rc, lc, &act, &nx);
        // does my transition lead exactly to the next lane's corner?
        const int cnext_lane = __shfl_down_sync(0xffffffffu, ci, 1);
        const bool open_next = __shfl_down_sync(0xffffffffu, (int)selfopen, 1) != 0;
        const bool trans = act == 0 && nx >= 0 && lane < 31 && nx == cnext_lane && open_next;
        const unsigned tmask = __ballot_sync(0xffffffffu, trans);
        const int m = __ffs(~tmask) - 1;                             // lanes 0..m execute (lane 31 never transitions)
        const bool exec = lane <= m;
        if (lane == 0 && !selfopen) status = UVOL_ERR_CORRUPT;        // the walk only ever moves to unvisited faces
        const unsigned newv = __ballot_sync(0xffffffffu, exec && !vis);
        if (n + __popc(newv) > max_entries) status = UVOL_ERR_CORRUPT;
        status = __shfl_sync(0xffffffffu, status, 0);
        if (status) break;
        if (exec) {
            if (G) fvis[fi] = 1; else atomicOr(&fbits[fi >> 5], 1u << (fi & 31));
            if (!vis) {
                const int idx = n + __popc(newv & lt);
                if (!G) atomicOr(&vbits[v >> 5], 1u << (v & 31));
                v2d1[v] = idx + 1; d2c[idx] = ci;
            }
        }
        n += __popc(newv);
        // outcome of the last executed lane
        const int act_m = __shfl_sync(0xffffffffu, act, m), nx_m = __shfl_sync(0xffffffffu, nx, m), lc_m = __shfl_sync(0xffffffffu, lc, m);
        int fm = f0 + m * pdir;
        if (pk != 32) fm = __shfl_sync(0xffffffffu, fi, m);
        __syncwarp();
        if (act_m == 0) { c = nx_m; if (c < 0) { status = UVOL_ERR_CORRUPT; break; } }
        else if (act_m == 1) { sp--; c = -1; }
        else {      // both neighbours open: the left face waits on the stack, the right one is walked next
            if (sp >= F + 4) { status = UVOL_ERR_CORRUPT; break; }
            if (lane == 0) {
                if (sp <= TRAV_STACK) stk[sp - 1] = lc_m; else gst[sp - 1] = lc_m;
                if (sp < TRAV_STACK) stk[sp] = nx_m; else gst[sp] = nx_m;
            }
            sp++; c = nx_m;
            __syncwarp();
        }
        // ---- the next step's guess (uniform): direction after a move to a neighbouring face id; after a jump, the row pattern if the
        // row that just ended and the distance between row starts repeat what the step before saw
        {
            int nk = 32, nD = 0;
            if (c >= 0) {
                const int nf = (int)face_of(c);
                if (nf == fm + 1) pdir = 1; else if (nf == fm - 1) pdir = -1;
                else {
                    int rowlen = m + 1, D = nf - f0;
                    if (pk != 32) { const int row0 = (m / pk) * pk; rowlen = m - row0 + 1; D = nf - __shfl_sync(0xffffffffu, fi, row0); }
                    const bool usable = rowlen <= 16 && (D > 32 || D < -32);
                    if (usable && ((rowlen == lastk && D == lastD) || (pk < 32 && rowlen == pk && D == pD))) { nk = rowlen; nD = D; }
                    lastk = rowlen; lastD = D;
                }
            } else { lastk = 0; lastD = 0; }
            if (pk < 32 && m == 31) { nk = pk; nD = pD; }          // a full step on the pattern: keep it
            if (nk != pk) { prow = nk == 32 ? 0 : lane / nk; pcol = lane - prow * nk; }
            pk = nk; pD = nD;
        }
    }
#undef FBIT
#undef VBIT
    if (lane == 0) {
        counts[jb.frame].entries[t] = (uint32_t)n;
        if (!status && (uint32_t)n != counts[jb.frame].expected[t]) status = UVOL_ERR_CORRUPT;    // the entropy runs were sized from `expected`
        if (status) frame_fail(counts, jb.frame, status);
    }
}

// Parallelogram parents, element-parallel.  grid = (ceil(maxN/128), frames, attrs)
// par4 = {opp entry, next entry, prev entry, kind}.  kind 1 ("scan-able"): the prediction is
// x[p-1] + (x[far1] - x[far2]) with both far parents at least 32 entries back -- then a run of such
// entries is a prefix sum (see k_predict_wrap); par4 is rewritten as {far1, far2, -, 1} (far = -1: term absent).
__global__ void __launch_bounds__(128) k_parents(const DracoFrame *frames, const DracoCounts *counts, uint8_t *S, const uint8_t *Z, const uint8_t *S2, const uint8_t *Z2) {
    const uint32_t fi = blockIdx.y, j = blockIdx.z;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if ((int)j >= f.nattr || f.o_par[j] == UVOL_NONE || (f.attr[j].pred != 1 && f.attr[j].pred != 0)) return;
    const int t = f.attr[j].table + 1, n = (int)counts[fi].entries[t];
    const TableView tv = make_view(f, t, S, Z);
    for (int p = blockIdx.x * 128 + threadIdx.x; p < n; p += gridDim.x * 128) {
        int par[4] = {-1, -1, -1, 0};
        if (f.attr[j].pred == 1) parallelogram_parents(p, tv, (const int *)(S2 + f.o_d2c[t]), (const int *)(Z2 + f.o_v2d[t]), par);
        int4 out = make_int4(par[0], par[1], par[2], 0);
        if (p > 0) {
            if (par[0] < 0) out = make_int4(-1, -1, -1, 1);                                              // delta coding: x[p-1] + corr
            else if (par[2] == p - 1 && par[0] <= p - 32 && par[1] <= p - 32) out = make_int4(par[1], par[0], -1, 1);   // + next - opp
            else if (par[1] == p - 1 && par[0] <= p - 32 && par[2] <= p - 32) out = make_int4(par[2], par[0], -1, 1);   // + prev - opp
        }
        ((int4 *)(S + f.o_par[j]))[p] = out;
    }
}

// DIFFERENCE / PARALLELOGRAM + WRAP reversal (also the pass-through for "no prediction"): one warp per
// (frame, attribute).  Runs of scan-able entries (k_parents) are reversed 32 at a time with a warp prefix
// sum; the wrap transform's clamp / wrap (rare: a handful per frame) is detected after the fact and the
// offending entry is redone with the exact serial formula, as are the entries of any other shape.
// Values of the last PW_RING entries are mirrored in a shared-memory ring (the far parents sit about one
// strip back).  what = attribute index.
#define PW_RING 512
__global__ void __launch_bounds__(32 * SERIAL_WARPS) k_predict_wrap(const DracoFrame *frames, const DracoCounts *counts, uint8_t *S, const Job *jobs, int njobs) {
    __shared__ int ring_all[SERIAL_WARPS][PW_RING * 4];
    const int ji = blockIdx.x * SERIAL_WARPS + (threadIdx.x >> 5);
    if (ji >= njobs) return;
    int *ring = ring_all[threadIdx.x >> 5];
    const Job jb = jobs[ji];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame]; const int j = jb.what; const DracoAttr &a = f.attr[j];
    const int n = (int)counts[jb.frame].entries[a.table + 1], lane = threadIdx.x & 31, nc = a.vnc;
    // values are reconstructed IN PLACE over the corrections: every entry's correction is read (by the lane that owns the entry)
    // before its value is stored, and only values of earlier entries are ever looked up
    const int32_t *corr = (const int32_t *)(S + f.o_corr[j]); int32_t *val = (int32_t *)(S + f.o_corr[j]);
    if (a.pred == -2 || n <= 0) return;                     // no prediction: the corrections are the values
    const int4 *par = (const int4 *)(S + f.o_par[j]);
    const int32_t mn = a.wmin, mx = a.wmax;
    // x(e, k): value of entry e, from the ring when recent enough
#define PW_GET(e, k, pcur) (((e) > (pcur) - PW_RING + 32) ? ring[((e) & (PW_RING - 1)) * 4 + (k)] : val[(e) * nc + (k)])
    int carry[4] = {0, 0, 0, 0};
    if (lane < nc) { const int32_t v = wrap_value(0, corr[lane], mn, mx); val[lane] = v; ring[lane] = v; }
    __syncwarp();
    for (int k = 0; k < nc; k++) carry[k] = ring[k];
    int p = 1;
    while (p < n) {
        const int q = p + lane;
        int4 pr = make_int4(-1, -1, -1, 0);
        if (q < n) pr = par[q];
        const unsigned amask = __ballot_sync(0xffffffffu, q < n && pr.w == 1);
        const int L = amask == 0xffffffffu ? 32 : __ffs(~amask) - 1;
        if (L > 0) {
            int g[4] = {0, 0, 0, 0}, cr[4] = {0, 0, 0, 0};
            if (lane < L) {
                for (int k = 0; k < nc; k++) {
                    cr[k] = corr[q * nc + k];
                    int far = 0;
                    if (pr.x >= 0) far = PW_GET(pr.x, k, p) - PW_GET(pr.y, k, p);
                    g[k] = far + cr[k];
                }
            }
            bool ok = true; int x[4];
            for (int k = 0; k < nc; k++) {
                int sc = g[k];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
                x[k] = carry[k] + sc;
                const int pred = x[k] - cr[k];
                ok = ok && pred >= mn && pred <= mx && x[k] >= mn && x[k] <= mx;
            }
            const unsigned okmask = __ballot_sync(0xffffffffu, ok || lane >= L);
            const int L2 = okmask == 0xffffffffu ? L : min(L, __ffs(~okmask) - 1);
            if (lane < L2) for (int k = 0; k < nc; k++) { val[q * nc + k] = x[k]; ring[(q & (PW_RING - 1)) * 4 + k] = x[k]; }
            if (L2 > 0) for (int k = 0; k < nc; k++) carry[k] = __shfl_sync(0xffffffffu, x[k], L2 - 1);
            p += L2;
            __syncwarp();
            if (L2 == L && L > 0) continue;
        }
        if (p >= n) break;
        // exact serial step for entry p (any shape, clamp and wrap applied): lane k owns component k
        const int4 ps = par[p];
        if (lane < nc) {
            long long pred;
            if (ps.w == 1) pred = (long long)carry[lane] + (ps.x >= 0 ? (long long)PW_GET(ps.x, lane, p) - PW_GET(ps.y, lane, p) : 0);
            else if (ps.x >= 0) pred = ((long long)PW_GET(ps.y, lane, p) + PW_GET(ps.z, lane, p)) - PW_GET(ps.x, lane, p);
            else pred = carry[lane];
            const int32_t v = wrap_value(pred, corr[p * nc + lane], mn, mx);
            val[p * nc + lane] = v; ring[(p & (PW_RING - 1)) * 4 + lane] = v;
        }
        __syncwarp();
        for (int k = 0; k < nc; k++) carry[k] = ring[(p & (PW_RING - 1)) * 4 + k];
        p += 1;
    }
#undef PW_GET
}

// TEX_COORDS_PORTABLE position-only terms, element-parallel.  grid = (ceil(maxN/128), frames, attrs)
__global__ void __launch_bounds__(128) k_uv_prepare(const DracoFrame *frames, const DracoCounts *counts, uint8_t *S, const uint8_t *Z, const uint8_t *S2, const uint8_t *Z2) {
    const uint32_t fi = blockIdx.y, j = blockIdx.z;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if ((int)j >= f.nattr || f.o_par[j] == UVOL_NONE || f.attr[j].pred != 5) return;
    const int t = f.attr[j].table + 1, n = (int)counts[fi].entries[t];
    const TableView tv = make_view(f, t, S, Z);
    const double vmax = fmax(fabs((double)f.attr[j].wmin), fabs((double)f.attr[j].wmax)) + 1.0, lim = 2251799813685248.0 /* 2^51 */;
    for (int p = blockIdx.x * 128 + threadIdx.x; p < n; p += gridDim.x * 128) {
        UvPrep q;
        uv_prepare(p, tv, (const int *)(S2 + f.o_d2c[t]), (const int *)(Z2 + f.o_v2d[t]), (const int *)(Z2 + f.o_v2d[0]),
                   (const int32_t *)(S + f.o_corr[f.pos_attr]), q);
        UvPrepD o; o.nd = q.nd; o.pd = q.pd; o.pn2 = o.dot = o.ns = o.rcp = 0.0;
        if (q.pn2 != 0) {
            const double dpn2 = (double)q.pn2, ddot = (double)q.dot, dns = (double)q.ns;
            if ((fabs(ddot) + dns) * 2.0 * vmax < lim && dpn2 * vmax < lim && dpn2 < lim) { o.pn2 = dpn2; o.dot = ddot; o.ns = dns; o.rcp = 1.0 / dpn2; }
            else { o.pn2 = __longlong_as_double(q.pn2); o.dot = __longlong_as_double(q.dot); o.ns = __longlong_as_double(q.ns); o.rcp = -1.0; }
        }
        ((UvPrepD *)(S + f.o_par[j]))[p] = o;
    }
}

// TEX_COORDS_PORTABLE chain: one warp per (frame, attribute).  The recurrence is non-linear (integer
// division by |PN|^2), so it stays a serial walk by lane 0, but everything that is not on the dependency
// chain is taken off it: the position-only terms and 1/|PN|^2 come from k_uv_prepare, tiles of 32 entries
// (terms, corrections, orientation flags) are staged in shared memory by the whole warp, the last UV_RING
// values are mirrored in a shared-memory ring, and the two truncating 64-bit divisions per entry are done as
// double-precision multiplies by the prepared reciprocal with an exact remainder fix-up.
#define UV_RING 1024
// trunc((n * pn2 + y) / pn2) for integers held exactly in doubles (|y|, |n * pn2| < 2^52): floor(y / pn2) by reciprocal
// multiply with a one-step exact remainder fix-up, then the toward-zero correction from the sign of the total.
__device__ __forceinline__ int32_t uv_div_fast(int32_t n, double y, double pn2, double rcp) {
    double qd = floor(y * rcp), rd = fma(-qd, pn2, y);
    if (rd < 0.0) { qd -= 1.0; rd += pn2; }
    if (rd >= pn2) { qd += 1.0; rd -= pn2; }
    const double total = fma((double)n, pn2, y);
    const int32_t adj = (total < 0.0 && rd != 0.0) ? 1 : 0;
    return (int32_t)((uint32_t)(unsigned long long)(long long)qd + (uint32_t)n + (uint32_t)adj);
}
struct UvTile { UvPrepD prep[32]; int32_t corr[64]; uint8_t orient[32]; int pad[8]; };
__global__ void __launch_bounds__(32 * SERIAL_WARPS) k_predict_uv(const DracoFrame *frames, DracoCounts *counts, uint8_t *S, const Job *jobs, int njobs) {
    __shared__ __align__(16) int ring_all[SERIAL_WARPS][UV_RING * 2];
    __shared__ UvTile tile_all[SERIAL_WARPS];
    const int ji = blockIdx.x * SERIAL_WARPS + (threadIdx.x >> 5);
    if (ji >= njobs) return;
    int *ring = ring_all[threadIdx.x >> 5]; UvTile &T = tile_all[threadIdx.x >> 5];
    const Job jb = jobs[ji];
    if (frame_dead(frames, counts, jb.frame)) return;
    const DracoFrame &f = frames[jb.frame]; const int j = jb.what; const DracoAttr &a = f.attr[j];
    const int n = (int)counts[jb.frame].entries[a.table + 1], lane = threadIdx.x & 31;
    const UvPrepD *prep = (const UvPrepD *)(S + f.o_par[j]); const int32_t *corr = (const int32_t *)(S + f.o_corr[j]);
    int32_t *uv = (int32_t *)(S + f.o_corr[j]); const uint8_t *orient = S + f.o_auxbits[j];          // in place: a tile's corrections are staged before its values are stored
    const int32_t mn = a.wmin, mx = a.wmax;
    int nor = a.num_orient, status = 0;
    for (int base = 0; base < n; base += 32) {
        // stage the tile (all lanes)
        const int cnt = min(32, n - base);
        if (lane < cnt) { T.prep[lane] = prep[base + lane]; T.corr[2 * lane] = corr[2 * (base + lane)]; T.corr[2 * lane + 1] = corr[2 * (base + lane) + 1]; }
        const int nor0 = __shfl_sync(0xffffffffu, nor, 0);                  // flags [nor0-32, nor0) cover this tile
        { const int oi = nor0 - 32 + lane; T.orient[lane] = oi >= 0 ? orient[oi] : 0; }
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < cnt; i++) {
                const int p = base + i; const UvPrepD q = T.prep[i];
                int pred0, pred1; bool have = false;
                if (q.pd < p && q.nd < p && q.pd >= 0 && q.nd >= 0) {
                    const int2 nv = q.nd > p - UV_RING ? ((const int2 *)ring)[q.nd & (UV_RING - 1)] : ((const int2 *)uv)[q.nd];
                    const int2 pv = q.pd > p - UV_RING ? ((const int2 *)ring)[q.pd & (UV_RING - 1)] : ((const int2 *)uv)[q.pd];
                    const int n0 = nv.x, n1 = nv.y, p0 = pv.x, p1 = pv.y;
                    if (n0 == p0 && n1 == p1) { pred0 = p0; pred1 = p1; have = true; }
                    else if (q.rcp != 0.0) {
                        if (nor <= 0) { status = UVOL_ERR_CORRUPT; break; }
                        --nor;
                        const bool o = T.orient[nor - (nor0 - 32)] != 0;
                        if (q.rcp > 0.0) {              // exact fp64: y = dot*d +- ns*d', pred = n + trunc-corrected floor(y / pn2)
                            const double dd0 = (double)(p0 - n0), dd1 = (double)(p1 - n1), sns = o ? q.ns : -q.ns;
                            const double y0 = fma(q.dot, dd0, sns * dd1), y1 = fma(q.dot, dd1, -(sns * dd0));
                            pred0 = uv_div_fast(n0, y0, q.pn2, q.rcp); pred1 = uv_div_fast(n1, y1, q.pn2, q.rcp);
                        } else {
                            const long long pn2 = __double_as_longlong(q.pn2), dot = __double_as_longlong(q.dot), ns = __double_as_longlong(q.ns);
                            const long long d0 = (long long)p0 - n0, d1 = (long long)p1 - n1;
                            const long long x0 = (long long)n0 * pn2 + dot * d0, x1 = (long long)n1 * pn2 + dot * d1;
                            const long long c0 = d1 * ns, c1 = -d0 * ns;
                            const long long a0 = o ? x0 + c0 : x0 - c0, a1 = o ? x1 + c1 : x1 - c1;
                            pred0 = (int32_t)(a0 / pn2); pred1 = (int32_t)(a1 / pn2);
                        }
                        have = true;
                    }
                }
                if (!have) {
                    if (q.nd < p && q.nd >= 0) { const int2 nv = q.nd > p - UV_RING ? ((const int2 *)ring)[q.nd & (UV_RING - 1)] : ((const int2 *)uv)[q.nd]; pred0 = nv.x; pred1 = nv.y; }
                    else if (p > 0) { const int2 lv = ((const int2 *)ring)[(p - 1) & (UV_RING - 1)]; pred0 = lv.x; pred1 = lv.y; }
                    else { pred0 = pred1 = 0; }
                }
                const int32_t u0 = wrap_value(pred0, T.corr[2 * i], mn, mx), u1 = wrap_value(pred1, T.corr[2 * i + 1], mn, mx);
                ((int2 *)ring)[p & (UV_RING - 1)] = make_int2(u0, u1);
                *(int2 *)(uv + 2 * p) = make_int2(u0, u1);
            }
        }
        __syncwarp();
        status = __shfl_sync(0xffffffffu, status, 0);
        if (status) break;
    }
    if (status && lane == 0) frame_fail(counts, jb.frame, status);
}

// GEOMETRIC_NORMAL, element-parallel (each entry depends only on finished positions); values in place over the corrections.
__global__ void __launch_bounds__(128) k_normals(const DracoFrame *frames, const DracoCounts *counts, uint8_t *S, const uint8_t *Z, const uint8_t *S2, const uint8_t *Z2) {
    const uint32_t fi = blockIdx.y, j = blockIdx.z;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    if ((int)j >= f.nattr || f.o_corr[j] == UVOL_NONE || f.attr[j].pred != 6) return;
    const int t = f.attr[j].table + 1, n = (int)counts[fi].entries[t];
    const TableView tv = make_view(f, t, S, Z);
    for (int p = blockIdx.x * 128 + threadIdx.x; p < n; p += gridDim.x * 128)
        normal_entry(p, tv, (const int *)(S2 + f.o_d2c[t]), (const int *)(Z2 + f.o_v2d[0]), (const int32_t *)(S + f.o_corr[f.pos_attr]),
                     (const int32_t *)(S + f.o_corr[j]), S + f.o_auxbits[j], f.attr[j].wmin, (int32_t *)(S + f.o_corr[j]));
}

// Per-point expansion + dequantisation into the output arrays (a5): one thread per point walks point -> corner once and then, per
// exported attribute, corner -> (attribute) vertex -> entry -> value -> fp32.  Output rows are written with streaming stores
// (nothing on the device reads them again).   grid = (ceil(capP/256), frames); the per-frame attribute constants sit in shared memory.
struct ExpandAttr { const int *voc, *v2d; const int32_t *val; float *out; int seq, nc, qbits, normalized, dtype; float qmin[4], qrange; };
__global__ void __launch_bounds__(256) k_expand(const DracoFrame *frames, const DracoCounts *counts, const uint8_t *S, const uint8_t *S2, const uint8_t *Z2, uint8_t *O, uint32_t slot_mask) {
    __shared__ ExpandAttr A[UVOL_MAX_ATTRS]; __shared__ int na;
    const uint32_t fi = blockIdx.y;
    if (frame_dead(frames, counts, fi)) return;
    const DracoFrame &f = frames[fi];
    const int P = (int)counts[fi].num_points;
    if ((int)(blockIdx.x * 256) >= P) return;
    if (threadIdx.x == 0) {
        int k = 0;
        for (int j = 0; j < f.nattr; j++) {
            const DracoAttr &a = f.attr[j]; if (a.out_slot < 0 || !((slot_mask >> a.out_slot) & 1u)) continue;
            const int t = a.table + 1; ExpandAttr &e = A[k++];
            e.voc = t == 0 ? (const int *)(S + f.o_c2v) : (const int *)(S + f.o_ac2v[t - 1]); e.v2d = (const int *)(Z2 + f.o_v2d[t]);
            e.val = (const int32_t *)(S + f.o_corr[j]); e.out = (float *)(O + f.out_attr[a.out_slot]);
            e.seq = a.seq; e.nc = a.nc; e.qbits = a.qbits; e.normalized = a.normalized; e.dtype = a.dtype; e.qrange = a.qrange;
            for (int c = 0; c < 4; c++) e.qmin[c] = a.qmin[c];
        }
        na = k;
    }
    __syncthreads();
    const int *p2c = (const int *)(S2 + f.o_p2c);
    for (int p = blockIdx.x * 256 + threadIdx.x; p < P; p += gridDim.x * 256) {
        const int c = p2c[p];
        for (int k = 0; k < na; k++) {
            const ExpandAttr &e = A[k];
            const int en = e.v2d[e.voc[c]] - 1; float *o = e.out + (size_t)p * e.nc;
            if (e.seq == 2) {
                const float delta = draco_dequant_delta(e.qrange, e.qbits);
                for (int q = 0; q < e.nc; q++) __stcs(o + q, draco_dequant(e.val[en * e.nc + q], delta, e.qmin[q]));
            } else if (e.seq == 3) {
                float r[3]; draco_oct_to_unit(e.val[en * 2], e.val[en * 2 + 1], ((1 << e.qbits) - 1) - 1, r);
                __stcs(o, r[0]); __stcs(o + 1, r[1]); __stcs(o + 2, r[2]);
            } else {
                const float tmax = e.dtype == 1 ? 127.f : e.dtype == 2 ? 255.f : e.dtype == 3 ? 32767.f : e.dtype == 4 ? 65535.f : e.dtype == 5 ? 2147483647.f : 4294967295.f;
                for (int q = 0; q < e.nc; q++) { float v = (float)e.val[en * e.nc + q]; if (e.normalized && e.dtype >= 1 && e.dtype <= 6) v = UVOL_FDIV(v, tmax); __stcs(o + q, v); }
            }
        }
    }
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

static const char *g_geo_stage_names[32] = {};
extern "C" const char *uvol_geo_stage_name(int i) { return (i >= 0 && i < 32 && g_geo_stage_names[i]) ? g_geo_stage_names[i] : ""; }

// Host-side state of the batch currently resident on the device (kept so that the device pipeline
// can be re-run on HBM-resident inputs, uvol_replay_draco_batch).
struct GeoBatch {
    std::vector<DracoFrame> frames; std::vector<uint32_t> aux; std::vector<Job> jobs;
    int n = 0, j_ransA = 0, j_rabsA = 0, j_trav = 0, j_ransB = 0, j_rabsB = 0, j_wrap = 0, j_uv = 0, j_end = 0;
    uint32_t max_alpha_ctx = 1, max_alpha_attr = 1, maxnad = 0, maxV = 0, maxF = 0; int maxattr = 0;
    uint64_t blob_bytes = 0, bytes_in = 0; DracoPlan pl; double parse_ms = 0; bool any_valence = false, any_standard = false, any_tagged = false;
    uint64_t cap_s2 = 0, cap_z2 = 0, cap_out = 0;        // what this batch may use of the count-sized arenas (estimates, or the exact needs after a re-plan)
    uint32_t replans = 0;
};
void uvol_geo_batch_free(GeoBatch *b) { delete b; }

// Test hooks (read per batch): shrink the optimistic per-frame attribute-table capacity / the estimates of the count-sized arenas so
// that ordinary meshes exercise the re-plan paths.
static uint32_t cap_permille() { const char *e = getenv("UVOL_CAP_PERMILLE"); const int x = e ? atoi(e) : 1000; return (uint32_t)(x < 1 ? 1 : x); }
static uint64_t est_scale(uint64_t v) { const char *e = getenv("UVOL_EST_PERMILLE"); const int x = e ? atoi(e) : 1000; return v * (uint64_t)(x < 1 ? 1 : x) / 1000; }

// Sizes every arena of the batch from the planner's numbers (grow-only reservations).
static int draco_reserve(uvol_ctx *ctx, int memory) {
    GeoBatch &B = *ctx->geo; const DracoPlan &pl = B.pl; const size_t n = (size_t)B.n;
    UVOL_CUDA(ctx, ctx->d_scratch.reserve(pl.scratch + 256));
    UVOL_CUDA(ctx, ctx->d_zscratch.reserve(pl.zscratch + 256));
    UVOL_CUDA(ctx, ctx->d_scratch2.reserve(B.cap_s2 + 256));
    UVOL_CUDA(ctx, ctx->d_zscratch2.reserve(B.cap_z2 + 256));
    UVOL_CUDA(ctx, ctx->d_out_geo.reserve(B.cap_out + 256));
    UVOL_CUDA(ctx, ctx->d_counts.reserve(align_up(sizeof(DracoCounts) * n, 16) + sizeof(DracoBatchPlan) + 256));
    UVOL_CUDA(ctx, ctx->h_counts.reserve(align_up(sizeof(DracoCounts) * n, 16) + sizeof(DracoBatchPlan) + 512 + sizeof(DracoFrame) * n));
    if (memory == UVOL_MEM_HOST) UVOL_CUDA(ctx, ctx->ph_out->reserve(B.cap_out + 256));
    return UVOL_OK;
}

static int draco_prepare(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n) {
    const double t_begin = now_ms();
    if (!ctx->geo) ctx->geo = new GeoBatch();
    GeoBatch &B = *ctx->geo;
    B.n = n; B.frames.assign((size_t)n, DracoFrame()); B.aux.clear(); B.aux.reserve((size_t)n * 2048); B.jobs.clear();
    B.max_alpha_ctx = B.max_alpha_attr = 1; B.maxnad = B.maxV = B.maxF = 0; B.maxattr = 0; B.bytes_in = 0; B.any_valence = B.any_standard = B.any_tagged = false; B.replans = 0;
    std::vector<DracoFrame> &frames = B.frames; std::vector<uint32_t> &aux = B.aux;
    uint64_t blob_bytes = 0;
    for (int i = 0; i < n; i++) {          // where every file goes in the blob follows from the sizes alone
        DracoFrame &f = frames[i]; memset(&f, 0, sizeof f);
        f.file_off = blob_bytes; f.file_len = (uint32_t)size[i]; B.bytes_in += size[i];
        blob_bytes = align_up(blob_bytes + size[i] + 8, 16);
    }
    B.blob_bytes = blob_bytes;
    UVOL_CUDA(ctx, ctx->h_blob.reserve(blob_bytes + 64));
    {   // Structural parse + staging copy into the pinned blob, both in front of the first kernel: a large batch is split over a few
        // threads, each parsing its frames into its own table area (spliced into `aux` afterwards, offsets rebased) and copying them.
        const int nthreads = std::max(1, std::min(n, (B.bytes_in > (4ull << 20) && n >= 16) ? (ctx->cfg.staging_threads ? (int)ctx->cfg.staging_threads : uvol_staging_threads()) : 1));
        const int per = (n + nthreads - 1) / nthreads;
        std::vector<std::vector<uint32_t>> laux((size_t)nthreads);
        auto work = [&](int t) {          // (uvol_draco_parse leaves file_off / file_len of the descriptor alone)
            std::vector<uint32_t> &la = nthreads == 1 ? aux : laux[(size_t)t];
            const int lo = t * per, hi = std::min(n, lo + per);
            if (nthreads > 1) la.reserve((size_t)std::max(0, hi - lo) * 2048);
            for (int i = lo; i < hi; i++) {
                const bool have = data[i] && size[i] < (1ull << 31);
                frames[i].status = have ? uvol_draco_parse(data[i], size[i], frames[i], la) : UVOL_ERR_ARG;
                if (have) memcpy((uint8_t *)ctx->h_blob.p + frames[i].file_off, data[i], size[i]);
            }
        };
        if (nthreads == 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; t++) pool.emplace_back(work, t);
            for (auto &t : pool) t.join();
            for (int t = 0; t < nthreads; t++) {          // splice the table areas, rebasing what points into them
                const uint32_t base = (uint32_t)aux.size();
                aux.insert(aux.end(), laux[(size_t)t].begin(), laux[(size_t)t].end());
                for (int i = t * per; i < std::min(n, (t + 1) * per); i++) {
                    DracoFrame &f = frames[i];
                    f.ts_off += base;
                    for (int k = 0; k < 6; k++) f.ctx[k].prob_off += base;
                    for (int j = 0; j < UVOL_MAX_ATTRS; j++) f.attr[j].sym.prob_off += base;
                }
            }
        }
    }
    for (int i = 0; i < n; i++) {
        DracoFrame &f = frames[i];
        // resource limits (per item, before anything is reserved): a header may not ask for more faces than the configured cap, nor
        // for absurdly more faces than the file has bytes (the densest real streams stay below one face per byte)
        if (!f.status && ((uint64_t)f.nf > ctx->cfg.max_faces_per_frame || (uint64_t)f.nf > 4096 + 64ull * size[i])) f.status = UVOL_ERR_UNSUPPORTED;
        if (f.status) continue;
        if (f.trav == 2) B.any_valence = true; else B.any_standard = true;
        for (int j = 0; j < f.nattr; j++) if (f.attr[j].tagged) B.any_tagged = true;
        for (int k = 0; k < 6; k++) if (f.ctx[k].count && f.ctx[k].nnz > B.max_alpha_ctx) B.max_alpha_ctx = f.ctx[k].nnz;      // (table sizes follow the used symbols)
        for (int j = 0; j < f.nattr; j++) {
            if (f.attr[j].sym.nnz > 8192) { f.status = UVOL_ERR_UNSUPPORTED; break; }
            if ((f.attr[j].out_slot >= 0 || j == f.pos_attr) && f.attr[j].sym.nnz > B.max_alpha_attr) B.max_alpha_attr = f.attr[j].sym.nnz;
        }
    }
    aux.push_back(0);
    draco_plan_phase1(frames, B.pl, cap_permille());
    B.cap_s2 = est_scale(B.pl.s2_est); B.cap_z2 = est_scale(B.pl.z2_est); B.cap_out = B.pl.out_index + est_scale(B.pl.out_est - B.pl.out_index);
    std::vector<Job> &jobs = B.jobs; jobs.reserve((size_t)n * 24);
    auto mark = [&]() { return (int)jobs.size(); };
    B.j_ransA = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int k = 0; k < 6; k++) if (frames[i].ctx[k].count) jobs.push_back({(uint32_t)i, k});
    B.j_rabsA = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (uint32_t k = 0; k < frames[i].nad; k++) jobs.push_back({(uint32_t)i, (int)k});
    B.j_trav = mark();
    for (int i = 0; i < n; i++) {
        const DracoFrame &f = frames[i]; if (f.status) continue;
        bool need[UVOL_MAX_ATTR_DATA + 1]; draco_tables_needed(f, need);
        for (uint32_t t = 0; t <= f.nad; t++) if (need[t]) jobs.push_back({(uint32_t)i, (int)t});
        if (f.nad > B.maxnad) B.maxnad = f.nad;
        if (f.nv_enc + f.nsplit > B.maxV) B.maxV = f.nv_enc + f.nsplit;
        if (f.nf > B.maxF) B.maxF = f.nf;
        if (f.nattr > B.maxattr) B.maxattr = f.nattr;
    }
    B.j_ransB = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int j = 0; j < frames[i].nattr; j++) if (draco_attr_needed(frames[i], j)) jobs.push_back({(uint32_t)i, 16 + j});
    B.j_rabsB = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int j = 0; j < frames[i].nattr; j++) if ((frames[i].attr[j].out_slot >= 0) && (frames[i].attr[j].pred == 5 || frames[i].attr[j].pred == 6)) jobs.push_back({(uint32_t)i, 16 + j});
    B.j_wrap = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int j = 0; j < frames[i].nattr; j++) { const DracoAttr &a = frames[i].attr[j]; if (draco_attr_needed(frames[i], j) && (a.pred == 0 || a.pred == 1)) jobs.push_back({(uint32_t)i, j}); }
    B.j_uv = mark();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int j = 0; j < frames[i].nattr; j++) if (frames[i].attr[j].out_slot >= 0 && frames[i].attr[j].pred == 5) jobs.push_back({(uint32_t)i, j});
    B.j_end = mark();
    // device buffers + uploads
    cudaStream_t st = ctx->s0;
    UVOL_CUDA(ctx, ctx->d_blob.reserve(blob_bytes + 64));
    UVOL_CUDA(ctx, ctx->d_desc.reserve(sizeof(DracoFrame) * (size_t)n));
    UVOL_CUDA(ctx, ctx->d_aux.reserve(aux.size() * 4));
    UVOL_CUDA(ctx, ctx->d_jobs.reserve(sizeof(Job) * (jobs.size() + 1)));
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

// Enqueues the whole device pipeline of the resident batch on the ctx's streams: no host synchronisation inside.  Ends with the
// small device -> host copy of the per-frame counts, the batch plan and the descriptors (with the offsets the device planner wrote).
static int draco_enqueue(uvol_ctx *ctx, int memory, bool first_attempt_of_fresh_upload, int *ev_out, int *i_seams_out, int *i_rans_out, int *i_rabs_out, uint32_t *launches_out, bool *attrs_copied) {
    *attrs_copied = false;
    GeoBatch &B = *ctx->geo; const int n = B.n;
    std::vector<DracoFrame> &frames = B.frames; DracoPlan &pl = B.pl;
    cudaStream_t st = ctx->s0;
    int ev = 1, i_seams = -1, i_rans = -1, i_rabs = -1;
    auto stamp = [&](const char *name) { if (ev < 32) g_geo_stage_names[ev - 1] = name; if (ctx->profile && ev < 32) cudaEventRecord(ctx->ev[ev], st); ev++; };
    if (!first_attempt_of_fresh_upload && ctx->profile) cudaEventRecord(ctx->ev[0], st);
    uint8_t *hd = (uint8_t *)ctx->h_desc.p;
    memcpy(hd, frames.data(), sizeof(DracoFrame) * (size_t)n);
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_desc.p, hd, sizeof(DracoFrame) * (size_t)n, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_counts.p, 0, align_up(sizeof(DracoCounts) * (size_t)n, 16) + sizeof(DracoBatchPlan), st));
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_zscratch.p, 0, pl.zscratch + 256, st));
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_zscratch2.p, 0, B.cap_z2 + 256, st));
    stamp("h2d");
    DracoFrame *dF = (DracoFrame *)ctx->d_desc.p; DracoCounts *dC = (DracoCounts *)ctx->d_counts.p; DracoBatchPlan *dBP = (DracoBatchPlan *)((uint8_t *)dC + align_up(sizeof(DracoCounts) * (size_t)n, 16));      // (DracoCounts is 108 bytes: the plan's 64-bit fields need their own alignment)
    const uint8_t *dBlob = (const uint8_t *)ctx->d_blob.p; const uint32_t *dAux = (const uint32_t *)ctx->d_aux.p;
    uint8_t *dS = (uint8_t *)ctx->d_scratch.p, *dZ = (uint8_t *)ctx->d_zscratch.p; const Job *dJ = (const Job *)ctx->d_jobs.p;
    uint8_t *dS2 = (uint8_t *)ctx->d_scratch2.p, *dZ2 = (uint8_t *)ctx->d_zscratch2.p, *dO = (uint8_t *)ctx->d_out_geo.p;
    uint32_t launches = 0;
    // context runs: 6-bit index, 4 per block; attribute runs: 8-bit index (2.5 KB + 16 B per used symbol) so that all runs of a
    // 1000-frame batch are resident at once
    const int lbA = B.max_alpha_ctx <= 8 ? 6 : RANS_LUT_BITS_MAX, wpbA = lbA == 6 ? 4 : 1, lbB = B.n > 256 ? 8 : RANS_LUT_BITS_MAX;
    auto rans_words = [](uint32_t nnz, int bits) { return (int)(((size_t)(nnz + 1) * 16 + (10u << bits) + 16 + 15) / 16 * 4); };
    auto rans_smem = [&](uint32_t alphabet, int bits, int wpb) { return (size_t)rans_words(alphabet, bits) * 4 * wpb; };
    auto nblk = [](int jobs) { return (unsigned)((jobs + SERIAL_WARPS - 1) / SERIAL_WARPS); };
    auto nblk32 = [](int jobs) { return (unsigned)((jobs + 31) / 32); };
    {
        const size_t smA = rans_smem(B.max_alpha_ctx, lbA, wpbA), smB = rans_smem(B.max_alpha_attr, lbB, 1);
        const size_t smMax = smA > smB ? smA : smB;
        if (smMax > 200 * 1024) { ctx->set_error("rANS alphabet too large for the shared-memory tables"); return UVOL_ERR_UNSUPPORTED; }
        if (smMax > 48 * 1024) UVOL_CUDA(ctx, cudaFuncSetAttribute(k_rans, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smMax));
        UVOL_CUDA(ctx, cudaFuncSetAttribute(k_edgebreaker_valence2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)SERIAL_WARPS * (EB2_RING + 128 + EB2_STAGE) * 16)));
    }
    // ---- phase 1.  The seam-bit runs depend only on the file bytes: they run on the side stream s1
    // next to the context-symbol runs and the connectivity walk.
    cudaStream_t sx = getenv("UVOL_NO_OVERLAP") ? ctx->s0 : ctx->s1;
    UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[0], st)); UVOL_CUDA(ctx, cudaStreamWaitEvent(sx, ctx->sync_ev[0], 0));
    if (ctx->profile) cudaEventRecord(ctx->aux_ev[0], sx);
    if (B.j_trav - B.j_rabsA > 0) { k_rabs_lanes<<<nblk32(B.j_trav - B.j_rabsA), 32, 0, sx>>>(dF, dC, dBlob, dS, dJ + B.j_rabsA, B.j_trav - B.j_rabsA); launches++; }
    if (ctx->profile) cudaEventRecord(ctx->aux_ev[1], sx);
    UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[1], sx));
    // The attribute symbol runs do not wait for the COUNTS the connectivity stages produce either: they run to the coder's terminal
    // state (k_rans, early mode) straight into the attribute's correction array, on the side stream behind the seam bits; a run
    // that did not come out with exactly the expected count is decoded again, by count, once the count is known.
    static const bool no_early = getenv("UVOL_NO_EARLY_RANS") != nullptr;
    if (ctx->profile) cudaEventRecord(ctx->aux_ev[2], sx);
    if (!no_early && B.j_rabsB - B.j_ransB > 0) { k_rans<<<nblk(B.j_rabsB - B.j_ransB), 32, rans_smem(B.max_alpha_attr, lbB, 1), sx>>>(dF, dC, dBlob, dAux, dS, dJ + B.j_ransB, B.j_rabsB - B.j_ransB, rans_words(B.max_alpha_attr, lbB), 1, lbB); launches++; }
    if (ctx->profile) cudaEventRecord(ctx->aux_ev[3], sx);
    UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[4], sx));
    if (B.j_rabsA - B.j_ransA > 0) { k_rans<<<(unsigned)((B.j_rabsA - B.j_ransA + wpbA - 1) / wpbA), 32 * wpbA, rans_smem(B.max_alpha_ctx, lbA, wpbA), st>>>(dF, dC, dBlob, dAux, dS, dJ + B.j_ransA, B.j_rabsA - B.j_ransA, rans_words(B.max_alpha_ctx, lbA), 0, lbA); launches++; }
    stamp("rans_ctx");
    stamp("rabs_seams(s1)"); i_seams = ev - 2;                                     // (stage slot of rabs_seams: timed on s1, filled in below)
    if (B.any_valence) {
        k_edgebreaker_valence2<<<nblk(n), 32 * SERIAL_WARPS, (size_t)SERIAL_WARPS * (EB2_RING + 128 + EB2_STAGE) * 16, st>>>(dF, dC, dBlob, dAux, dS, n);
        launches++;
    }
    if (B.any_standard) { k_edgebreaker<<<nblk(n), 32 * SERIAL_WARPS, 0, st>>>(dF, dC, dBlob, dAux, dS, n); launches++; }
    stamp("edgebreaker");
    UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[9], st));
    // The base table's traversal records need the corner table only: they are built on the side stream next to the seam /
    // attribute-table / point stages (their arena lies clear of those stages' temporaries, draco_plan.h).
    const int ntj_all = B.j_ransB - B.j_trav;
    const bool early_base = pl.base_records_early && ntj_all > 0 && !getenv("UVOL_NO_OVERLAP");
    if (early_base) {
        UVOL_CUDA(ctx, cudaStreamWaitEvent(ctx->s3, ctx->sync_ev[9], 0));      // (s1 is busy with the attribute symbol runs)
        k_face_records<<<dim3((B.maxF + 255) / 256, ntj_all), 256, 0, ctx->s3>>>(dF, dC, dS, dZ, dJ + B.j_trav, 0); launches++;
        UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[10], ctx->s3));
    }
    UVOL_CUDA(ctx, cudaStreamWaitEvent(st, ctx->sync_ev[1], 0));
    if (B.maxnad) {
        const dim3 sg((3 * B.maxF + SEAM_CHUNK - 1) / SEAM_CHUNK, (unsigned)n);
        k_seam_count<<<sg, 256, 0, st>>>(dF, dC, dS); k_seams<<<sg, 256, 0, st>>>(dF, dC, dS, dZ, n); launches += 2;
    }
    stamp("seams");
    const unsigned gv = (B.maxV + 127) / 128 > 0 ? (B.maxV + 127) / 128 : 1;
    if (B.maxnad) {
        k_attr_fan<0><<<dim3(gv, n, B.maxnad), 128, 0, st>>>(dF, dC, dS, dZ); launches++;
        k_scan<<<dim3(n, B.maxnad), 1024, 0, st>>>(dF, dC, dS, 0); launches++;
        k_attr_fan<1><<<dim3(gv, n, B.maxnad), 128, 0, st>>>(dF, dC, dS, dZ); launches++;
    }
    stamp("attr_tables");
    k_point_fan<0><<<dim3(gv, n), 128, 0, st>>>(dF, dC, dS, dZ, nullptr, nullptr); launches++;
    k_scan<<<dim3(n, 1), 1024, 0, st>>>(dF, dC, dS, 4); launches++;
    stamp("point_count");
    // ---- the count-sized arrays are laid out on the device; everything below is enqueued without waiting for it
    k_plan2<<<1, PLAN_T, 0, st>>>(dF, dC, dBP, n, pl.out_index, B.cap_s2, B.cap_z2, B.cap_out); launches++;
    stamp("plan2");
    if (memory == UVOL_MEM_HOST) {      // the batch plan (region bases of the output arena) travels to the host now: it is read below, once everything is enqueued
        UVOL_CUDA(ctx, ctx->h_aux.reserve(256));
        UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_aux.p, dBP, sizeof(DracoBatchPlan), cudaMemcpyDeviceToHost, st));
        UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[5], st));
    }
    // ---- phase 2.  The aux bit runs (sized from counts.expected) go to s1 and overlap the traversals.
    UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[2], st)); UVOL_CUDA(ctx, cudaStreamWaitEvent(sx, ctx->sync_ev[2], 0));
    if (ctx->profile) cudaEventRecord(ctx->aux_ev[5], sx);
    if (B.j_wrap - B.j_rabsB > 0) { k_rabs_lanes<<<nblk32(B.j_wrap - B.j_rabsB), 32, 0, sx>>>(dF, dC, dBlob, dS, dJ + B.j_rabsB, B.j_wrap - B.j_rabsB); launches++; }
    if (ctx->profile) cudaEventRecord(ctx->aux_ev[4], sx);
    UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[3], sx));
    k_point_fan<1><<<dim3(gv, n), 128, 0, st>>>(dF, dC, dS, dZ, dS2, dO); launches++;
    stamp("point_assign");
    if (memory == UVOL_MEM_HOST && pl.out_index) {      // the index buffers are final: copy them out on s3 while the rest of phase 2 runs
        UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[6], st)); UVOL_CUDA(ctx, cudaStreamWaitEvent(ctx->s3, ctx->sync_ev[6], 0));
        UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->ph_out->p, dO, pl.out_index, cudaMemcpyDeviceToHost, ctx->s3));
    }
    if (B.j_ransB - B.j_trav > 0) {
        const int ntj = B.j_ransB - B.j_trav;
        k_face_records<<<dim3((B.maxF + 255) / 256, ntj), 256, 0, st>>>(dF, dC, dS, dZ, dJ + B.j_trav, early_base ? 1 : 2); launches++;
        if (early_base) UVOL_CUDA(ctx, cudaStreamWaitEvent(st, ctx->sync_ev[10], 0));
        stamp("face_records");
        uint32_t capN = B.maxV + 4;
        for (int i = 0; i < n; i++) if (!frames[i].status) for (uint32_t t = 1; t <= frames[i].nad; t++) if (frames[i].table_cap[t] > capN) capN = frames[i].table_cap[t];
        const int fwords = (int)(((B.maxF + 31) / 32 + 4) & ~3u), vwords = (int)(((capN + 31) / 32 + 4) & ~3u);
        const size_t smem = ((size_t)(fwords + vwords) * 4 + TRAV_STACK * 4);
        // mode 0: both bitmaps in shared memory -- fastest while the walks of the batch need at most ~3 waves of the SMs' shared memory;
        // mode 1: byte map + vertex -> entry map in global memory, every walk resident at once (large meshes / large batches)
        static const int force = getenv("UVOL_TRAV_GMAP") ? atoi(getenv("UVOL_TRAV_GMAP")) : -1;
        const size_t per_sm = (size_t)((ntj + ctx->num_sms - 1) / ctx->num_sms), fit = smem <= 200 * 1024 ? (200 * 1024) / smem : 0;
        int mode = force >= 0 ? (force != 0) : (fit > 0 && per_sm <= 3 * fit ? 0 : 1);
        if (mode == 0 && fit == 0) mode = 1;
        if (mode == 1) k_traverse<1><<<ntj, 32, TRAV_STACK * 4, st>>>(dF, dC, dS, dZ, dS2, dZ2, dJ + B.j_trav, ntj, fwords, vwords);
        else {
            if (smem > 48 * 1024) UVOL_CUDA(ctx, cudaFuncSetAttribute(k_traverse<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_traverse<0><<<ntj, 32, smem, st>>>(dF, dC, dS, dZ, dS2, dZ2, dJ + B.j_trav, ntj, fwords, vwords);
        }
        launches++;
    } else stamp("face_records");
    stamp("traverse");
    UVOL_CUDA(ctx, cudaStreamWaitEvent(st, ctx->sync_ev[4], 0));
    if (B.j_rabsB - B.j_ransB > 0) {      // attribute runs whose early decode did not end on the expected count (normally none): again, by count
        const int nrj = B.j_rabsB - B.j_ransB;
        k_rans<<<nblk(nrj), 32, rans_smem(B.max_alpha_attr, lbB, 1), st>>>(dF, dC, dBlob, dAux, dS, dJ + B.j_ransB, nrj, rans_words(B.max_alpha_attr, lbB), 0, lbB); launches++;
    }
    if (B.any_tagged) { k_tagged_values<<<dim3((unsigned)n, (unsigned)B.maxattr), 256, 0, st>>>(dF, dC, dBlob, dS); launches++; }
    stamp("rans_recheck");
    UVOL_CUDA(ctx, cudaStreamWaitEvent(st, ctx->sync_ev[3], 0));
    stamp("rans_attr(s1)"); i_rans = ev - 2;                                     // (stage slots of rans_attr / rabs_aux: timed on s1)
    stamp("rabs_aux(s1)"); i_rabs = ev - 2;
    const unsigned gn = (pl.cap_entries + 127) / 128;
    k_parents<<<dim3(gn, n, B.maxattr), 128, 0, st>>>(dF, dC, dS, dZ, dS2, dZ2); launches++;
    stamp("parents");
    if (B.j_uv - B.j_wrap > 0) { k_predict_wrap<<<nblk(B.j_uv - B.j_wrap), 32 * SERIAL_WARPS, 0, st>>>(dF, dC, dS, dJ + B.j_wrap, B.j_uv - B.j_wrap); launches++; }
    stamp("predict_wrap");
    k_normals<<<dim3(gn, n, B.maxattr), 128, 0, st>>>(dF, dC, dS, dZ, dS2, dZ2); launches++;
    stamp("normals");
    // positions, normals and colours are final: expand them now so that their spans of the output arena can travel to the host
    // while the UV chain still runs
    const dim3 gexp((pl.cap_points + 255) / 256, n);
    k_expand<<<gexp, 256, 0, st>>>(dF, dC, dS, dS2, dZ2, dO, 0xbu); launches++;
    stamp("expand_pnc");
    UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[7], st));
    k_uv_prepare<<<dim3(gn, n, B.maxattr), 128, 0, st>>>(dF, dC, dS, dZ, dS2, dZ2); launches++;
    stamp("uv_prepare");
    if (B.j_end - B.j_uv > 0) { k_predict_uv<<<nblk(B.j_end - B.j_uv), 32 * SERIAL_WARPS, 0, st>>>(dF, dC, dS, dJ + B.j_uv, B.j_end - B.j_uv); launches++; }
    stamp("predict_uv");
    k_expand<<<gexp, 256, 0, st>>>(dF, dC, dS, dS2, dZ2, dO, 0x4u); launches++;
    stamp("expand");
    UVOL_CUDA(ctx, cudaEventRecord(ctx->sync_ev[8], st));
    ctx->span_geo_end = ev - 1 < 32 ? ev - 1 : 31;
    uint8_t *hC = (uint8_t *)ctx->h_counts.p; const size_t cbytes = align_up(sizeof(DracoCounts) * (size_t)n, 16) + sizeof(DracoBatchPlan);
    UVOL_CUDA(ctx, cudaMemcpyAsync(hC, dC, cbytes, cudaMemcpyDeviceToHost, st));
    UVOL_CUDA(ctx, cudaMemcpyAsync(hC + align_up(cbytes, 256), dF, sizeof(DracoFrame) * (size_t)n, cudaMemcpyDeviceToHost, st));
    *ev_out = ev; *i_seams_out = i_seams; *i_rans_out = i_rans; *i_rabs_out = i_rabs; *launches_out = launches;
    if (memory == UVOL_MEM_HOST) {
        // Everything is enqueued; the stream keeps running while this thread waits for the plan (ready once the connectivity stages
        // are done) and then queues the result copies on s3 behind the kernels that finish each span.
        UVOL_CUDA(ctx, cudaEventSynchronize(ctx->sync_ev[5]));
        const DracoBatchPlan plan = *(const DracoBatchPlan *)ctx->h_aux.p;
        if (!plan.overflow && plan.out_need <= B.cap_out && plan.out_need > pl.out_index) {
            uint8_t *hO = (uint8_t *)ctx->ph_out->p;
            const uint64_t first = plan.slot_base[0], mid = plan.slot_base[2], end = plan.out_need;       // [position | normal | colour] then [uv]
            UVOL_CUDA(ctx, cudaStreamWaitEvent(ctx->s3, ctx->sync_ev[7], 0));
            if (mid > first) UVOL_CUDA(ctx, cudaMemcpyAsync(hO + first, dO + first, mid - first, cudaMemcpyDeviceToHost, ctx->s3));
            UVOL_CUDA(ctx, cudaStreamWaitEvent(ctx->s3, ctx->sync_ev[8], 0));
            if (end > mid) UVOL_CUDA(ctx, cudaMemcpyAsync(hO + mid, dO + mid, end - mid, cudaMemcpyDeviceToHost, ctx->s3));
            *attrs_copied = true;
        }
    }
    return UVOL_OK;
}

static int draco_run(uvol_ctx *ctx, int memory, uvol_geometry *out, bool fresh_upload) {
    GeoBatch &B = *ctx->geo; const int n = B.n;
    std::vector<DracoFrame> &frames = B.frames; DracoPlan &pl = B.pl;
    cudaStream_t st = ctx->s0;
    int ev = 1, i_seams = -1, i_rans = -1, i_rabs = -1; uint32_t launches = 0, launches_total = 0; bool attrs_copied = false;
    const size_t cbytes = align_up(sizeof(DracoCounts) * (size_t)n, 16) + sizeof(DracoBatchPlan);
    const DracoCounts *hC = nullptr; const DracoBatchPlan *hBP = nullptr; const DracoFrame *hF = nullptr;
    for (int attempt = 0;; attempt++) {
        draco_plan_phase1(frames, pl, cap_permille());      // (re)establishes the header-sized layout (idempotent; honours full_cap)
        int rc = draco_reserve(ctx, memory); if (rc) return rc;
        rc = draco_enqueue(ctx, memory, fresh_upload && attempt == 0, &ev, &i_seams, &i_rans, &i_rabs, &launches, &attrs_copied); if (rc) return rc;
        launches_total += launches;
        UVOL_CUDA(ctx, cudaStreamSynchronize(st));
        hC = (const DracoCounts *)ctx->h_counts.p; hBP = (const DracoBatchPlan *)((const uint8_t *)ctx->h_counts.p + align_up(sizeof(DracoCounts) * (size_t)n, 16)); hF = (const DracoFrame *)((const uint8_t *)ctx->h_counts.p + align_up(cbytes, 256));
        // Re-plan and run again when the optimistic sizing did not hold (at most twice: the second plan is exact / uses the full bounds).
        bool again = false;
        if (hBP->overflow) { B.cap_s2 = std::max(B.cap_s2, hBP->s2_need); B.cap_z2 = std::max(B.cap_z2, hBP->z2_need); B.cap_out = std::max(B.cap_out, hBP->out_need); again = true; }
        for (int i = 0; i < n; i++) if (!frames[i].status && hC[i].status == UVOL_ERR_FRAME_CAPACITY && !frames[i].full_cap) { frames[i].full_cap = 1; again = true; }
        if (!again || attempt >= 2) break;
        B.replans++;
        if (memory == UVOL_MEM_HOST) UVOL_CUDA(ctx, cudaStreamSynchronize(ctx->s3));
    }
    pl.scratch2 = hBP->s2_need; pl.zscratch2 = hBP->z2_need; pl.out = hBP->out_need;
    uint8_t *dO = (uint8_t *)ctx->d_out_geo.p;
    if (memory == UVOL_MEM_HOST && !hBP->overflow && pl.out > pl.out_index && !attrs_copied)
        UVOL_CUDA(ctx, cudaMemcpyAsync((uint8_t *)ctx->ph_out->p + pl.out_index, dO + pl.out_index, pl.out - pl.out_index, cudaMemcpyDeviceToHost, st));
    { if (ev < 32) g_geo_stage_names[ev - 1] = "d2h"; if (ctx->profile && ev < 32) cudaEventRecord(ctx->ev[ev], st); ev++; }
    UVOL_CUDA(ctx, cudaStreamSynchronize(st));
    if (memory == UVOL_MEM_HOST) UVOL_CUDA(ctx, cudaStreamSynchronize(ctx->s3));
    UVOL_CUDA(ctx, cudaGetLastError());
    // ---- results (offsets of the count-sized arrays: as written by the device planner)
    uint8_t *base = memory == UVOL_MEM_HOST ? (uint8_t *)ctx->ph_out->p : dO;
    uint64_t bytes_out = 0;
    for (int i = 0; i < n; i++) {
        const DracoFrame &f = frames[i], &df = hF[i]; uvol_geometry &g = out[i];
        memset(&g, 0, sizeof g);
        g.status = f.status ? f.status : hC[i].status;
        if (g.status <= UVOL_ERR_FRAME_CAPACITY) g.status = UVOL_ERR_UNSUPPORTED;      // internal codes never leave the library
        if (g.status) continue;
        g.num_points = hC[i].num_points; g.num_faces = f.nf;
        g.index = (uint32_t *)(base + f.out_index); bytes_out += (uint64_t)f.nf * 12;
        for (int j = 0; j < f.nattr; j++) {
            const DracoAttr &a = f.attr[j]; if (a.out_slot < 0) continue;
            float *p = (float *)(base + df.out_attr[a.out_slot]); bytes_out += (uint64_t)g.num_points * a.nc * 4;
            if (a.out_slot == 0) g.position = p; else if (a.out_slot == 1) g.normal = p; else if (a.out_slot == 2) g.uv = p; else { g.color = p; g.color_components = (uint32_t)a.nc; }
        }
    }
    uvol_stats &s = ctx->stats;
    s.kernel_launches = launches_total; s.bytes_in = B.bytes_in; s.bytes_out = bytes_out;
    s.scratch_bytes = pl.scratch + pl.zscratch + pl.scratch2 + pl.zscratch2;
    if (ctx->profile) {
        s.num_stages = (uint32_t)(ev - 1);
        for (int k = 0; k + 1 < ev && k < 24; k++) cudaEventElapsedTime(&s.stage_ms[k], ctx->ev[k], ctx->ev[k + 1]);
        float tot = 0; cudaEventElapsedTime(&tot, ctx->ev[1], ctx->ev[ev - 2]); s.device_ms = tot;   // kernels only: after h2d, before d2h
        s.h2d_ms = s.stage_ms[0]; s.d2h_ms = s.stage_ms[ev - 2];
        // stages that ran on the side stream (overlapped with the main stream): their own event pairs
        if (i_seams >= 0 && i_seams < 24) cudaEventElapsedTime(&s.stage_ms[i_seams], ctx->aux_ev[0], ctx->aux_ev[1]);
        if (i_rans >= 0 && i_rans < 24) cudaEventElapsedTime(&s.stage_ms[i_rans], ctx->aux_ev[2], ctx->aux_ev[3]);
        if (i_rabs >= 0 && i_rabs < 24) cudaEventElapsedTime(&s.stage_ms[i_rabs], ctx->aux_ev[5], ctx->aux_ev[4]);
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

// used by the combined V2 entry point (basis_transcode.cu)
int uvol_geo_prepare_and_run(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int memory, uvol_geometry *out, bool replay) {
    if (replay) { if (!ctx->geo || ctx->geo->n != n) return UVOL_ERR_ARG; return draco_run(ctx, memory, out, false); }
    int rc = draco_prepare(ctx, data, size, n); if (rc) return rc;
    rc = draco_run(ctx, memory, out, true); if (rc) return rc;
    ctx->stats.host_parse_ms = ctx->geo->parse_ms;
    return UVOL_OK;
}
