// corto_decode.cu -- sm_100a kernels + launcher for the V1 geometry path (Corto .crt frames), and the
// reference's own C ABI (CreateDecoder / DecodeMesh / DestroyDecoder, deprecated/encoder/dev/src/corto_codec.h:41-43).
//
// Replaces CortoDecoder.decode (src/lib/corto.ts:73-140) == crt::Decoder::decode
// (deprecated/encoder/dev/src/decoder.cpp:122-173) for a batch of frames sliced out of a .drcs
// (src/V1/worker.ts:48-68).  Layout: SURVEY.md Appendix C.  Stages:
//   tunstall   one warp per Tunstall block: lane 0 rebuilds the 256-word dictionary (tunstall.cpp:125-256),
//              then the warp expands the code bytes with a prefix sum over word lengths (:430-452)
//   values     one CTA per (frame, attribute): prefix sum of the per-value bit widths -> every value's bit
//              offset -> parallel bit extraction (cstream.h:296-362; bitstream.cpp:103-121, MSB-first words)
//   faces      one CTA per frame: warp 0 runs the front-growing connectivity walk (corto_core.h; decoder.cpp:181-333), one more warp
//              per attribute follows it with the parallelogram / delta reversal (vertex_attribute.h:155-177; normals with DIFF
//              prediction: normal_attribute.cpp:182-204) of the vertices whose context the walk has published
//   estimate   (normals with ESTIMATED / BORDER prediction, normal_attribute.cpp:24-59,206-303): vertex -> incident corners lists
//              (count, scan, fill), then per vertex the face normals are summed IN FACE ORDER (the reference accumulates floats
//              face by face, so the order is part of the result) and the boundary mark is XOR-ed together
//   dequant    element-parallel: (float)value * q (vertex_attribute.h:179-224); octahedral -> unit normals
//              (normal_attribute.h:104-112); YCC -> RGB colours times the per-channel step (color_attribute.cpp:69-90, point.h:214)
// Attributes: "position" (3 x f32), "uv" (2 x f32), "normal" (codec 2, all three predictions, 3 x f32), "color" (codec 3, RGBA8);
// other generic attributes are skipped over (they are self-delimiting).  Tunstall or no entropy coding.
#include <chrono>
#include <string.h>
#include <string>
#include <vector>
#include <thread>
#include "uvol_ctx.h"
#include "corto_core.h"
#include "corto_parse.h"
#include "../../include/corto_codec.h"

namespace {

// ------------------------------------------------------------------------------------------------- kernels
#define CW 4   // warps per block for the serial kernels

// Tunstall dictionary (tunstall.cpp:125-256, wordsize 8) + expansion.  what: 0 = clers, 1 + 4*attr + k = logs block k
struct TunSmem { uint32_t queues[512], index[512], lengths[512], starts[256]; uint8_t table[8192 + 64]; uint8_t sym[256], prob[256]; };
__global__ void __launch_bounds__(32 * CW) k_tunstall(const CortoFrame *frames, int32_t *status, const uint8_t *blob, uint8_t *S, const CJob *jobs, int njobs) {
    extern __shared__ uint8_t tun_smem[];
    const int ji = blockIdx.x * CW + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ji >= njobs) return;
    TunSmem &T = *(TunSmem *)(tun_smem + (size_t)(threadIdx.x >> 5) * sizeof(TunSmem));
    const CJob jb = jobs[ji];
    if (frames[jb.frame].status || status[jb.frame]) return;
    const CortoFrame &f = frames[jb.frame];
    const TunBlock &tb = jb.what == 0 ? f.clers : f.attr[(jb.what - 1) >> 2].logs[(jb.what - 1) & 3];
    const uint8_t *file = blob + f.file_off, *src = file + tb.data_off; uint8_t *out = S + tb.o_out;
    if (tb.size == 0) return;
    if (tb.nsym == 0xffffffffu) { for (uint32_t i = lane; i < tb.size; i += 32) out[i] = src[i]; return; }
    const uint32_t ns = tb.nsym;
    if (ns == 0) { if (lane == 0) status[jb.frame] = UVOL_ERR_CORRUPT; return; }
    for (uint32_t i = lane; i < ns; i += 32) { T.sym[i] = file[tb.probs_off + 2 * i]; T.prob[i] = file[tb.probs_off + 2 * i + 1]; }
    __syncwarp();
    if (ns == 1) { for (uint32_t i = lane; i < tb.size; i += 32) out[i] = T.sym[0]; return; }
    int bad = 0;
    if (lane == 0) {
        const uint32_t dict = 256; uint32_t end = 0, pos = 0, n_words = 0, count = 2;
        const uint32_t p0 = (uint32_t)T.prob[0] << 8, p1 = (uint32_t)T.prob[1] << 8; uint32_t prob = (p0 * p0) >> 16;
        const uint32_t max_count = (dict - 1) / (ns - 1);
        while (prob > p1 && count < max_count) { prob = (prob * p0) >> 16; count++; }
        if (count >= 16) {          // very low entropy: A..A, A..AB, A..AC words (tunstall.cpp:155-196)
            T.table[pos++] = T.sym[0];
            for (uint32_t k = 1; k < ns; k++) { for (uint32_t i = 0; i < count - 1; i++) T.table[pos++] = T.sym[0]; T.table[pos++] = T.sym[k]; }
            T.starts[0] = (count - 1) * ns; for (uint32_t k = 1; k < ns; k++) T.starts[k] = k;
            for (uint32_t col = 0; col < count; col++) {
                for (uint32_t row = 1; row < ns; row++) {
                    const uint32_t dest = row + col * ns;
                    if (dest >= 512) { bad = 1; break; }
                    T.queues[dest] = col == 0 ? ((uint32_t)T.prob[row] << 8) : ((prob * ((uint32_t)T.prob[row] << 8)) >> 16);
                    T.index[dest] = row * count - col; T.lengths[dest] = col + 1;
                }
                if (bad) break;
                prob = col == 0 ? p0 : (prob * p0) >> 16;
            }
            const uint32_t first = (count - 1) * ns;
            if (first >= 512) bad = 1;
            else { T.queues[first] = prob; T.index[first] = 0; T.lengths[first] = count; }
            n_words = 1 + count * (ns - 1); end = count * ns;
        } else {
            n_words = ns;
            for (uint32_t i = 0; i < ns; i++) { T.starts[i] = i; T.queues[end] = (uint32_t)T.prob[i] << 8; T.index[end] = pos; T.lengths[end++] = 1; T.table[pos++] = T.sym[i]; }
        }
        while (!bad && n_words < dict) {
            uint32_t best = 0, max_prob = 0;
            for (uint32_t i = 0; i < ns; i++) { const uint32_t p = T.queues[T.starts[i]]; if (p > max_prob) { best = i; max_prob = p; } }
            const uint32_t symbol = T.starts[best], probability = T.queues[symbol], offset = T.index[symbol], length = T.lengths[symbol];
            uint32_t r = 0;
            for (; r < ns; r++) {
                if (end >= 512 || pos + length + 1 > 8192) { bad = 1; break; }
                T.queues[end] = (probability * ((uint32_t)T.prob[r] << 8)) >> 16; T.index[end] = pos; T.lengths[end++] = length + 1;
                for (uint32_t k = 0; k < length; k++) T.table[pos + k] = T.table[offset + k];
                pos += length; T.table[pos++] = T.sym[r];
                if (n_words + r == dict - 1) break;
            }
            if (r == ns) T.starts[best] += ns;
            n_words += ns - 1;
        }
        uint32_t word = 0;
        for (uint32_t i = 0, row = 0; i < end && !bad; i++, row++) {        // compact index / lengths
            if (row >= ns) row = 0;
            if (T.starts[row] > i) continue;
            T.index[word] = T.index[i]; T.lengths[word] = T.lengths[i]; word++;
        }
        if (word < dict) for (uint32_t i = word; i < dict; i++) { T.index[i] = 0; T.lengths[i] = 0; }
    }
    __syncwarp();
    bad = __shfl_sync(0xffffffffu, bad, 0);
    if (bad || tb.csize == 0) { if (lane == 0) status[jb.frame] = UVOL_ERR_CORRUPT; return; }
    // expansion: every code byte but the last copies its whole word; the last one fills what is left (:430-452)
    uint32_t base = 0;
    for (uint32_t i0 = 0; i0 < tb.csize; i0 += 32) {
        const uint32_t i = i0 + lane; uint32_t len = 0, start = 0;
        if (i < tb.csize) { const uint32_t s = src[i]; len = T.lengths[s]; start = T.index[s]; }
        uint32_t inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        uint32_t off = base + inc - len;
        if (i < tb.csize) {
            if (i == tb.csize - 1) len = off < tb.size ? tb.size - off : 0;
            if (off + len > tb.size) len = off < tb.size ? tb.size - off : 0;
            if (start + len > sizeof(T.table)) len = start < sizeof(T.table) ? (uint32_t)sizeof(T.table) - start : 0;
            for (uint32_t k = 0; k < len; k++) out[off + k] = T.table[start + k];
        }
        base += __shfl_sync(0xffffffffu, inc, 31);
    }
}

__device__ __forceinline__ int ilog2_u(uint32_t p) { int k = 0; while (p >>= 1) ++k; return k; }

// Connectivity + delta reversal, one block per frame.
//   warp 0, lane 0: the front-growing walk (corto_core.h).  One walk per block, so the walks of a batch spread over all SMs.
//   warp 1 + a, lanes 0..N-1: the delta reversal of attribute a (value = correction + prediction from the parallelogram context the
//     walk recorded for the vertex, deltaDecode, vertex_attribute.h:154-189: with the PARALLEL strategy values[i] += values[a] +
//     values[b] - values[c], else += values[a]; normals with DIFF prediction always use the single parent, normal_attribute.cpp:
//     182-204; lane k owns component k; unsigned arithmetic: colours are decoded modulo 256 -- their C++ type is uchar -- and sums
//     modulo 2^32 reduce to the same residues).  The warps FOLLOW the walk: it publishes how many vertices have their context
//     (every 16 vertices, behind a block-level fence), they reverse what is published -- the chain costs about half of what the
//     walk costs per vertex, so it is hidden entirely instead of running as a second serial stage after the walk.
//   Both chains are latency chains, so what they touch again soon stays in shared memory: the walk's last CORTO_RING front records,
//   each delta warp's last CORTO_VRING values (a parent is, as a rule, a vertex decoded just before; older ones are read from the array).
#define CORTO_RING 1024      // front records kept in shared memory per walk (16 KB)
#define CORTO_VRING 512      // values kept in shared memory per attribute (8 KB)
#define CORTO_FACES_SMEM ((size_t)(CORTO_RING + CORTO_MAX_ATTRS * CORTO_VRING) * 16)
__global__ void __launch_bounds__(32 * (1 + CORTO_MAX_ATTRS)) k_corto_faces(const CortoFrame *frames, int32_t *status, const uint8_t *blob, const uint32_t *aux, uint8_t *S, uint8_t *O, int nframes) {
    extern __shared__ uint4 corto_smem[];
    __shared__ volatile int progress, walk_done;
    const int fi = blockIdx.x, warp = threadIdx.x >> 5, k = threadIdx.x & 31;
    if (fi >= nframes) return;
    if (frames[fi].status) { if (threadIdx.x == 0) status[fi] = frames[fi].status; return; }
    if (status[fi]) return;
    const CortoFrame &f = frames[fi];
    if (threadIdx.x == 0) { progress = 0; walk_done = 0; }
    __syncthreads();
    if (warp == 0) {
        if (k != 0) return;
        CortoWalkMem m;
        m.clers = S + f.clers.o_out; m.nclers = f.clers.size;
        m.bits = CortoBits{(const uint32_t *)(blob + f.file_off + f.ibits.data_off), 0, (uint64_t)f.ibits.nwords * 32};
        m.group_end = aux + f.groups_off; m.ngroups = f.ngroups;
        m.front = (CortoEdge *)(S + f.o_front); m.third = (uint32_t *)(S + f.o_third); m.front_cap = 3 * (int)f.nface + 8;
        m.ring = (CortoEdge *)corto_smem; m.ring_size = CORTO_RING;
        m.queue = (int *)(S + f.o_queue); m.delayed = (int *)(S + f.o_delayed); m.order_cap = 3 * (int)f.nface + 8;
        m.faces = (uint32_t *)(O + f.out_index); m.pred = (int *)(S + f.o_pred); m.nvert = (int)f.nvert; m.nface = (int)f.nface;
        m.progress = &progress;
        const int rc = corto_walk(m);
        if (rc) status[fi] = rc == CORTO_TRUNCATED ? UVOL_ERR_TRUNCATED : UVOL_ERR_CORRUPT;
        __threadfence_block();
        if (!rc) progress = (int)f.nvert;
        __threadfence_block();
        walk_done = 1;
        return;
    }
    // ---- delta reversal of attribute warp - 1
    const int ai = warp - 1;
    if (ai >= f.nattr) return;
    const CortoAttr &a = f.attr[ai];
    if ((a.kind == CK_NORMAL && a.pred != 0) || k >= a.vN || k >= 4) return;
    uint32_t *ring = (uint32_t *)(corto_smem + CORTO_RING + (size_t)ai * CORTO_VRING);
    uint32_t *val = (uint32_t *)(S + a.o_val); const int4 *pred = (const int4 *)(S + f.o_pred);
    const int n = (int)f.nvert, N = a.vN; const bool par = a.kind != CK_NORMAL && (a.strategy & 1) != 0;
    if (n > 0) ring[k] = val[k];
#define CORTO_PARENT(j) ((j) >= i - CORTO_VRING ? ring[((j) & (CORTO_VRING - 1)) * 4 + k] : val[(j) * N + k])
    int seen = 0;
    for (int i = 1; i < n;) {
        while (seen <= i) {                                   // vertex i has its context once progress > i
            seen = progress;
            if (seen > i) break;
            if (walk_done) { seen = progress; if (seen <= i) return; break; }      // the walk ended short of this vertex: a corrupt stream, its status is set
            __nanosleep(200);
        }
        const int hi = seen < n ? seen : n;
        int4 p = __ldcg(pred + i), p2 = i + 1 < hi ? __ldcg(pred + i + 1) : p;       // (L2 reads: the walk is writing this array)
        uint32_t c = val[i * N + k], c2 = i + 1 < hi ? val[(i + 1) * N + k] : 0u;
        for (; i < hi; i++) {
            const int4 p3 = i + 2 < hi ? __ldcg(pred + i + 2) : p2;       // contexts and corrections are requested two vertices ahead
            const uint32_t c3 = i + 2 < hi ? val[(i + 2) * N + k] : 0u;
            uint32_t v = c;
            if (par) v += CORTO_PARENT(p.x) + CORTO_PARENT(p.y) - CORTO_PARENT(p.z); else v += CORTO_PARENT(p.x);
            ring[(i & (CORTO_VRING - 1)) * 4 + k] = v;
            val[i * N + k] = v;
            p = p2; p2 = p3; c = c2; c2 = c3;
        }
    }
#undef CORTO_PARENT
}

// Per-value bit widths -> bit offsets -> values (cstream.h:296-362).  grid = (frames, attrs), 256 threads.
__global__ void __launch_bounds__(256) k_corto_values(const CortoFrame *frames, const int32_t *status, const uint8_t *blob, uint8_t *S) {
    __shared__ unsigned long long wsum[8], carry_s;
    const uint32_t fi = blockIdx.x, ai = blockIdx.y;
    if (frames[fi].status || status[fi]) return;
    const CortoFrame &f = frames[fi];
    if ((int)ai >= f.nattr) return;
    const CortoAttr &a = f.attr[ai];
    // normals always travel as one correlated array of two (normal_attribute.cpp:161-175), colours per component
    const uint32_t *words = (const uint32_t *)(blob + f.file_off + a.bits.data_off);
    int32_t *val = (int32_t *)(S + a.o_val);
    const int n = (int)a.count, N = a.vN, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool corr = a.kind == CK_NORMAL ? true : (a.kind == CK_COLOR ? false : (a.strategy & 2) != 0);
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int c = 0; c < (corr ? 1 : N); c++) {
        const uint8_t *logs = S + a.logs[c].o_out;
        for (int base = 0; base < n; base += 256) {
            const int i = base + tid; const int d = i < n ? logs[i] : 0;
            const unsigned long long mine = (unsigned long long)(corr ? d * N : d);
            unsigned long long inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) wsum[w] = inc;
            __syncthreads();
            unsigned long long pre = carry_s;
            for (int k = 0; k < w; k++) pre += wsum[k];
            unsigned long long bp = pre + inc - mine;
            if (i < n) {
                if (d > 25) { /* widths above 25 bits cannot come from a sane quantisation */ }
                if (corr) {
                    const int mx = d ? (1 << d) >> 1 : 0;
                    for (int k = 0; k < N; k++) {
                        int v = 0;
                        if (d) { const uint64_t wi = bp >> 5; const int sh = (int)(bp & 31); const uint64_t two = ((uint64_t)words[wi] << 32) | words[wi + 1]; v = (int)((two << sh) >> (64 - d)) - mx; bp += d; }
                        val[i * N + k] = v;
                    }
                } else {
                    int v = 0;
                    if (d) { const uint64_t wi = bp >> 5; const int sh = (int)(bp & 31); const uint64_t two = ((uint64_t)words[wi] << 32) | words[wi + 1]; v = (int)((two << sh) >> (64 - d)); const int middle = 1 << (d - 1); if (v < middle) v = -v - middle; }
                    val[i * N + c] = v;
                }
            }
            __syncthreads();
            if (tid == 255) carry_s = pre + inc;
            __syncthreads();
        }
    }
}

// Prediction reversal (vertex_attribute.h:155-177): values[i] += values[a] + values[b] - values[c] (PARALLEL)
// or += values[a]; normals with DIFF prediction always use the single parent (normal_attribute.cpp:182-204).  One warp per
// (frame, attribute), lane k owns component k.  Unsigned arithmetic: colours are decoded modulo 256 (their C++ type is uchar), and
// sums modulo 2^32 reduce to the same residues.
// ---- normal estimation (ESTIMATED / BORDER prediction)
// corners per vertex.  grid = (ceil(3 * maxF / 256), frames)
__global__ void __launch_bounds__(256) k_corto_vcount(const CortoFrame *frames, const int32_t *status, uint8_t *S, const uint8_t *O) {
    const uint32_t fi = blockIdx.y;
    if (frames[fi].status || status[fi]) return;
    const CortoFrame &f = frames[fi];
    if (f.nrm_attr < 0 || f.attr[f.nrm_attr].pred == 0) return;
    const uint32_t c = blockIdx.x * 256 + threadIdx.x;
    if (c >= 3 * f.nface) return;
    const uint32_t v = ((const uint32_t *)(O + f.out_index))[c];
    if (v < f.nvert) atomicAdd((int *)(S + f.o_voff) + v, 1);
}
// In-place exclusive scan of an int array of nvert (+1 total slot) entries.  which: 0 = corner counts, 1 = boundary flags.  grid = frames, 1024 threads
__global__ void __launch_bounds__(1024) k_corto_scan(const CortoFrame *frames, const int32_t *status, uint8_t *S, int which) {
    __shared__ int wsum[32]; __shared__ int carry_s;
    const uint32_t fi = blockIdx.x;
    if (frames[fi].status || status[fi]) return;
    const CortoFrame &f = frames[fi];
    if (f.nrm_attr < 0 || f.attr[f.nrm_attr].pred == 0 || (which == 1 && f.attr[f.nrm_attr].pred != 2)) return;
    int *a = (int *)(S + (which ? f.o_bflag : f.o_voff));
    const int V = (int)f.nvert, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
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
    if (tid == 0) a[V] = carry_s;
}
__global__ void __launch_bounds__(256) k_corto_vfill(const CortoFrame *frames, const int32_t *status, uint8_t *S, const uint8_t *O) {
    const uint32_t fi = blockIdx.y;
    if (frames[fi].status || status[fi]) return;
    const CortoFrame &f = frames[fi];
    if (f.nrm_attr < 0 || f.attr[f.nrm_attr].pred == 0) return;
    const uint32_t c = blockIdx.x * 256 + threadIdx.x;
    if (c >= 3 * f.nface) return;
    const uint32_t v = ((const uint32_t *)(O + f.out_index))[c];
    if (v >= f.nvert) return;
    const int slot = atomicAdd((int *)(S + f.o_vfill) + v, 1);
    ((uint32_t *)(S + f.o_vlist))[((const int *)(S + f.o_voff))[v] + slot] = c;
}
// Per vertex: incident corners sorted by corner id (= face order), face normals of the INTEGER positions summed in that order
// (estimateNormals, normal_attribute.cpp:40-59: float cross products accumulated face by face), boundary mark = XOR of the other
// two vertex ids of every incident corner (markBoundary, :24-37).  grid = (ceil(maxV / 128), frames)
__global__ void __launch_bounds__(128) k_corto_estimate(const CortoFrame *frames, const int32_t *status, uint8_t *S, const uint8_t *O) {
    const uint32_t fi = blockIdx.y;
    if (frames[fi].status || status[fi]) return;
    const CortoFrame &f = frames[fi];
    if (f.nrm_attr < 0 || f.attr[f.nrm_attr].pred == 0) return;
    const uint32_t v = blockIdx.x * 128 + threadIdx.x;
    if (v >= f.nvert) return;
    const int *off = (const int *)(S + f.o_voff); uint32_t *list = (uint32_t *)(S + f.o_vlist) + off[v]; const int n = off[v + 1] - off[v];
    for (int i = 1; i < n; i++) { const uint32_t x = list[i]; int j = i - 1; while (j >= 0 && list[j] > x) { list[j + 1] = list[j]; j--; } list[j + 1] = x; }
    const uint32_t *faces = (const uint32_t *)(O + f.out_index); const int32_t *pos = (const int32_t *)(S + f.attr[f.pos_attr].o_val);
    float ex = 0.f, ey = 0.f, ez = 0.f; int bmark = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t c = list[i], fb = c - c % 3u, k = c - fb;
        const uint32_t i0 = faces[fb], i1 = faces[fb + 1], i2 = faces[fb + 2];
        if (i0 >= f.nvert || i1 >= f.nvert || i2 >= f.nvert) continue;
        const float p0x = (float)pos[3 * i0], p0y = (float)pos[3 * i0 + 1], p0z = (float)pos[3 * i0 + 2];
        const float ax = __fsub_rn((float)pos[3 * i1], p0x), ay = __fsub_rn((float)pos[3 * i1 + 1], p0y), az = __fsub_rn((float)pos[3 * i1 + 2], p0z);
        const float bx = __fsub_rn((float)pos[3 * i2], p0x), by = __fsub_rn((float)pos[3 * i2 + 1], p0y), bz = __fsub_rn((float)pos[3 * i2 + 2], p0z);
        ex = __fadd_rn(ex, __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by)));
        ey = __fadd_rn(ey, __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz)));
        ez = __fadd_rn(ez, __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx)));
        bmark ^= k == 0 ? (int)(i1 ^ i2) : (k == 1 ? (int)(i2 ^ i0) : (int)(i0 ^ i1));
    }
    float *est = (float *)(S + f.o_est) + 3 * (size_t)v; est[0] = ex; est[1] = ey; est[2] = ez;
    ((int *)(S + f.o_bflag))[v] = bmark != 0;
}

// Octahedral mapping of the reference (normal_attribute.h:75-112), every float operation rounded on its own like the reference build
__device__ __forceinline__ void corto_to_octa(float x, float y, float z, int unit, int *ox, int *oy) {
    const float len = __fadd_rn(__fadd_rn(fabsf(x), fabsf(y)), fabsf(z));
    float px = __fdiv_rn(x, len), py = __fdiv_rn(y, len);
    if (z < 0.f) { const float qx = __fsub_rn(1.0f, fabsf(py)), qy = __fsub_rn(1.0f, fabsf(px)); px = x < 0.f ? -qx : qx; py = y < 0.f ? -qy : qy; }
    *ox = (int)__fmul_rn(px, (float)unit); *oy = (int)__fmul_rn(py, (float)unit);
}
__device__ __forceinline__ void corto_to_sphere(int vx, int vy, int unit, float *o) {
    float nx = (float)vx, ny = (float)vy, nz = (float)(unit - abs(vx) - abs(vy));
    if (nz < 0.f) { nx = (float)((vx > 0 ? 1 : -1) * (unit - abs(vy))); ny = (float)((vy > 0 ? 1 : -1) * (unit - abs(vx))); }
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
    const float len = (float)__dsqrt_rn((double)s);
    o[0] = __fdiv_rn(nx, len); o[1] = __fdiv_rn(ny, len); o[2] = __fdiv_rn(nz, len);
}

// Dequantisation of every attribute.  grid = (ceil(maxV / 256), frames, attrs)
//   position / uv : (float)value * q, one rounding (vertex_attribute.h:186-187)
//   normal DIFF   : toSphere(value) (normal_attribute.cpp:228-232); ESTIMATED / BORDER : computeNormals (:278-303)
//   colour        : YCC -> RGB, times the per-channel step, RGBA8 (color_attribute.cpp:69-90, point.h:214)
__global__ void __launch_bounds__(256) k_corto_dequant(const CortoFrame *frames, const int32_t *status, const uint8_t *S, uint8_t *O) {
    const uint32_t fi = blockIdx.y, ai = blockIdx.z;
    if (frames[fi].status || status[fi]) return;
    const CortoFrame &f = frames[fi];
    if ((int)ai >= f.nattr) return;
    const CortoAttr &a = f.attr[ai];
    const uint32_t v = blockIdx.x * 256 + threadIdx.x;
    if (v >= f.nvert) return;
    const int32_t *val = (const int32_t *)(S + a.o_val);
    if (a.kind == CK_POSITION || a.kind == CK_UV) {
        float *o = (float *)(O + a.out) + (size_t)v * a.N;
        for (int k = 0; k < a.N; k++) o[k] = __fmul_rn((float)val[v * a.N + k], a.q);
    } else if (a.kind == CK_NORMAL) {
        float *o = (float *)(O + a.out) + 3 * (size_t)v; const int unit = (int)a.q;
        if (a.pred == 0) { corto_to_sphere(val[2 * v], val[2 * v + 1], unit, o); return; }
        const float *e = (const float *)(S + f.o_est) + 3 * (size_t)v; const int *bf = (const int *)(S + f.o_bflag);
        // BORDER: corrections exist only for boundary vertices, in vertex order (bflag has been scanned: bf[v + 1] - bf[v] is the mark)
        const bool corrected = a.pred == 1 || bf[v + 1] != bf[v];
        if (corrected) {
            const uint32_t ci = a.pred == 1 ? v : (uint32_t)bf[v];
            int qx, qy; corto_to_octa(e[0], e[1], e[2], unit, &qx, &qy);
            const int dx = ci < a.count ? val[2 * ci] : 0, dy = ci < a.count ? val[2 * ci + 1] : 0;
            corto_to_sphere(qx + dx, qy + dy, unit, o);
        } else {
            const float s = __fadd_rn(__fadd_rn(__fmul_rn(e[0], e[0]), __fmul_rn(e[1], e[1])), __fmul_rn(e[2], e[2]));
            const float len = (float)__dsqrt_rn((double)s);
            o[0] = __fdiv_rn(e[0], len); o[1] = __fdiv_rn(e[1], len); o[2] = __fdiv_rn(e[2], len);
        }
    } else {
        uint32_t c[4] = {0, 0, 0, 255};
        for (int k = 0; k < a.N; k++) c[k] = (uint32_t)val[v * a.N + k] & 255u;
        const uint32_t rgb[4] = {(c[2] + c[0]) & 255u, c[0], (c[1] + c[0]) & 255u, c[3]};
        uint8_t *o = O + a.out + 4 * (size_t)v;
        for (int k = 0; k < 4; k++) o[k] = (uint8_t)(rgb[k] * a.qc[k]);
    }
}

// u32 -> u16 index narrowing (the web player's layout: src/V1/player.ts:292, corto.ts:675-680).  grid = (ceil(3 * maxF / 256), frames)
__global__ void __launch_bounds__(256) k_corto_index16(const CortoFrame *frames, const int32_t *status, uint8_t *O) {
    const uint32_t fi = blockIdx.y;
    if (frames[fi].status || status[fi] || !frames[fi].index16) return;
    const CortoFrame &f = frames[fi];
    const uint32_t c = blockIdx.x * 256 + threadIdx.x;
    if (c < 3 * f.nface) ((uint16_t *)(O + f.out_index16))[c] = (uint16_t)((const uint32_t *)(O + f.out_index))[c];
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
uint64_t take(uint64_t &cur, uint64_t bytes) { uint64_t o = cur; cur = (cur + bytes + 127) / 128 * 128; return o; }

}  // namespace

struct CortoBatch { std::vector<CortoFrame> frames; };
void uvol_corto_batch_free(CortoBatch *b) { delete b; }
static const char *kCortoStages[] = {"h2d", "tunstall", "values", "faces", "estimate", "dequant", "d2h"};
extern "C" const char *uvol_corto_stage_name(int i) { return (i >= 0 && i < 7) ? kCortoStages[i] : ""; }

extern "C" int uvol_decode_corto_batch(uvol_ctx *ctx, const uint8_t *const *data, const size_t *size, int n, int memory, uvol_corto_mesh *out) {
    if (!ctx || !out || n < 0 || (n > 0 && (!data || !size))) return UVOL_ERR_ARG;
    UVOL_CUDA(ctx, cudaSetDevice(ctx->device));
    memset(&ctx->stats, 0, sizeof ctx->stats);
    for (int i = 0; i < n; i++) memset(&out[i], 0, sizeof out[i]);
    if (n == 0) return UVOL_OK;
    const double t0 = now_ms();
    if (!ctx->corto) ctx->corto = new CortoBatch();
    std::vector<CortoFrame> &frames = ctx->corto->frames; frames.assign((size_t)n, CortoFrame());
    std::vector<uint32_t> aux; std::vector<CJob> jobs;
    uint64_t blob_bytes = 0, s = 0, z = 0, o = 0; uint32_t maxV = 1, maxF = 1; uint64_t bytes_in = 0; bool any_est = false, any16 = false;
    for (int i = 0; i < n; i++) {
        CortoFrame &f = frames[i]; memset(&f, 0, sizeof f);
        f.file_off = blob_bytes; f.file_len = (uint32_t)size[i]; bytes_in += size[i];
        blob_bytes = align_up(blob_bytes + size[i] + 16, 16);      // 4-byte alignment of the bit streams is preserved (decoder.cpp:42-43)
        f.status = (data[i] && size[i] < (1ull << 31)) ? corto_parse(data[i], size[i], f, aux, ctx->cfg.max_faces_per_frame) : UVOL_ERR_ARG;
        if (f.status) continue;
        f.clers.o_out = take(s, (uint64_t)f.clers.size + 16);
        const uint64_t cap = 3ull * f.nface + 8;
        f.o_front = take(s, cap * sizeof(CortoEdge)); f.o_third = take(s, cap * 4); f.o_queue = take(s, cap * 4); f.o_delayed = take(s, cap * 4);
        f.o_pred = take(s, ((uint64_t)f.nvert + 2) * 16);
        f.out_index = take(o, (uint64_t)f.nface * 12);
        f.index16 = ctx->cfg.corto_index_u16 && f.nface < 65536;
        if (f.index16) { f.out_index16 = take(o, (uint64_t)f.nface * 6); any16 = true; }
        if (f.nvert > maxV) maxV = f.nvert;
        if (f.nface > maxF) maxF = f.nface;
        for (int a = 0; a < f.nattr; a++) {
            CortoAttr &at = f.attr[a];
            for (int k = 0; k < at.nlogs; k++) at.logs[k].o_out = take(s, (uint64_t)at.logs[k].size + 8);
            at.o_val = take(s, ((uint64_t)f.nvert + 1) * at.vN * 4);
            at.out = take(o, at.kind == CK_COLOR ? (uint64_t)f.nvert * 4 : (uint64_t)f.nvert * (at.kind == CK_NORMAL ? 3 : at.N) * 4);
        }
        if (f.nrm_attr >= 0 && f.attr[f.nrm_attr].pred != 0) {      // zero-initialised: corner counts / fill cursors / boundary flags
            any_est = true;
            f.o_voff = take(z, ((uint64_t)f.nvert + 2) * 4); f.o_vfill = take(z, ((uint64_t)f.nvert + 2) * 4); f.o_bflag = take(z, ((uint64_t)f.nvert + 2) * 4);
            f.o_vlist = take(s, 3ull * f.nface * 4 + 16); f.o_est = take(s, (uint64_t)f.nvert * 12 + 16);
        }
    }
    // the zero-initialised part follows the uninitialised scratch in one arena
    const uint64_t zbase = align_up(s, 256);
    for (int i = 0; i < n; i++) if (!frames[i].status && frames[i].nrm_attr >= 0 && frames[i].attr[frames[i].nrm_attr].pred != 0) { frames[i].o_voff += zbase; frames[i].o_vfill += zbase; frames[i].o_bflag += zbase; }
    aux.push_back(0);
    const int j_tun = 0;
    for (int i = 0; i < n; i++) if (!frames[i].status) { jobs.push_back({(uint32_t)i, 0}); for (int a = 0; a < frames[i].nattr; a++) for (int k = 0; k < frames[i].attr[a].nlogs; k++) jobs.push_back({(uint32_t)i, 1 + 4 * a + k}); }
    const int j_end = (int)jobs.size();
    const size_t desc_bytes = sizeof(CortoFrame) * (size_t)n, aux_bytes = aux.size() * 4, job_bytes = sizeof(CJob) * (jobs.size() + 1);
    UVOL_CUDA(ctx, ctx->h_cblob.reserve(blob_bytes + 64)); UVOL_CUDA(ctx, ctx->d_cblob.reserve(blob_bytes + 64));
    UVOL_CUDA(ctx, ctx->h_cdesc.reserve(desc_bytes + aux_bytes + job_bytes + 64)); UVOL_CUDA(ctx, ctx->d_cdesc.reserve(desc_bytes + aux_bytes + job_bytes + 64));
    UVOL_CUDA(ctx, ctx->d_cscratch.reserve(zbase + z + 256)); UVOL_CUDA(ctx, ctx->d_out_corto.reserve(o + 256));
    UVOL_CUDA(ctx, ctx->d_ccounts.reserve(4 * (size_t)n)); UVOL_CUDA(ctx, ctx->h_ccounts.reserve(4 * (size_t)n));
    {   // staging copy into the pinned blob (it sits in front of the first kernel): only the padding behind each file is zeroed, and a
        // large batch is copied by a few threads
        uint8_t *hb = (uint8_t *)ctx->h_cblob.p;
        auto copy_range = [&](int lo, int hi) {
            for (int i = lo; i < hi; i++) {
                const bool have = data[i] && size[i] < (1ull << 31);
                const uint64_t end = i + 1 < n ? frames[i + 1].file_off : blob_bytes + 64, used = have ? size[i] : 0;
                if (have) memcpy(hb + frames[i].file_off, data[i], size[i]);
                memset(hb + frames[i].file_off + used, 0, end - frames[i].file_off - used);
            }
        };
        const int nthreads = bytes_in > (16ull << 20) ? (ctx->cfg.staging_threads ? (int)ctx->cfg.staging_threads : uvol_staging_threads()) : 1;
        if (nthreads <= 1) copy_range(0, n);
        else {
            std::vector<std::thread> pool; const int per = (n + nthreads - 1) / nthreads;
            for (int lo = 0; lo < n; lo += per) pool.emplace_back(copy_range, lo, std::min(n, lo + per));
            for (auto &t : pool) t.join();
        }
    }
    uint8_t *hd = (uint8_t *)ctx->h_cdesc.p;
    memcpy(hd, frames.data(), desc_bytes); memcpy(hd + desc_bytes, aux.data(), aux_bytes); memcpy(hd + desc_bytes + aux_bytes, jobs.data(), sizeof(CJob) * jobs.size());
    const double t_parsed = now_ms();
    cudaStream_t st = ctx->s0; int ev = 0;
    auto stamp = [&]() { if (ctx->profile && ev < 32) cudaEventRecord(ctx->ev[ev], st); ev++; };
    stamp();
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_cblob.p, ctx->h_cblob.p, blob_bytes + 64, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_cdesc.p, hd, desc_bytes + aux_bytes + job_bytes, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_ccounts.p, 0, 4 * (size_t)n, st));
    if (z) UVOL_CUDA(ctx, cudaMemsetAsync((uint8_t *)ctx->d_cscratch.p + zbase, 0, z, st));
    stamp();
    const CortoFrame *dF = (const CortoFrame *)ctx->d_cdesc.p; const uint32_t *dAux = (const uint32_t *)((const uint8_t *)ctx->d_cdesc.p + desc_bytes);
    const CJob *dJ = (const CJob *)((const uint8_t *)ctx->d_cdesc.p + desc_bytes + aux_bytes);
    int32_t *dSt = (int32_t *)ctx->d_ccounts.p; const uint8_t *dBlob = (const uint8_t *)ctx->d_cblob.p;
    uint8_t *dS = (uint8_t *)ctx->d_cscratch.p, *dO = (uint8_t *)ctx->d_out_corto.p;
    uint32_t launches = 0;
    auto nb = [](int j) { return (unsigned)((j + CW - 1) / CW); };
    if (j_end - j_tun > 0) {
        UVOL_CUDA(ctx, cudaFuncSetAttribute(k_tunstall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(TunSmem) * CW)));
        k_tunstall<<<nb(j_end - j_tun), 32 * CW, sizeof(TunSmem) * CW, st>>>(dF, dSt, dBlob, dS, dJ + j_tun, j_end - j_tun); launches++;
    }
    stamp();
    k_corto_values<<<dim3(n, CORTO_MAX_ATTRS), 256, 0, st>>>(dF, dSt, dBlob, dS); launches++;
    stamp();
    UVOL_CUDA(ctx, cudaFuncSetAttribute(k_corto_faces, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CORTO_FACES_SMEM));
    k_corto_faces<<<n, 32 * (1 + CORTO_MAX_ATTRS), CORTO_FACES_SMEM, st>>>(dF, dSt, dBlob, dAux, dS, dO, n); launches++;
    stamp();
    if (any_est) {
        const dim3 gc((3 * maxF + 255) / 256, n);
        k_corto_vcount<<<gc, 256, 0, st>>>(dF, dSt, dS, dO); k_corto_scan<<<n, 1024, 0, st>>>(dF, dSt, dS, 0); k_corto_vfill<<<gc, 256, 0, st>>>(dF, dSt, dS, dO);
        k_corto_estimate<<<dim3((maxV + 127) / 128, n), 128, 0, st>>>(dF, dSt, dS, dO); k_corto_scan<<<n, 1024, 0, st>>>(dF, dSt, dS, 1); launches += 5;
    }
    stamp();
    k_corto_dequant<<<dim3((maxV + 255) / 256, n, CORTO_MAX_ATTRS), 256, 0, st>>>(dF, dSt, dS, dO); launches++;
    if (any16) { k_corto_index16<<<dim3((3 * maxF + 255) / 256, n), 256, 0, st>>>(dF, dSt, dO); launches++; }
    stamp();
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_ccounts.p, dSt, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (memory == UVOL_MEM_HOST) { UVOL_CUDA(ctx, ctx->h_cout.reserve(o + 256)); UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->h_cout.p, dO, o, cudaMemcpyDeviceToHost, st)); }
    stamp();
    UVOL_CUDA(ctx, cudaStreamSynchronize(st));
    UVOL_CUDA(ctx, cudaGetLastError());
    const int32_t *hSt = (const int32_t *)ctx->h_ccounts.p; uint8_t *base = memory == UVOL_MEM_HOST ? (uint8_t *)ctx->h_cout.p : dO; uint64_t bytes_out = 0;
    for (int i = 0; i < n; i++) {
        const CortoFrame &f = frames[i]; uvol_corto_mesh &m = out[i];
        m.status = f.status ? f.status : hSt[i];
        if (m.status) continue;
        m.num_vertices = f.nvert; m.num_faces = f.nface; m.index = (uint32_t *)(base + f.out_index); bytes_out += (uint64_t)f.nface * 12;
        if (f.index16) { m.index16 = (uint16_t *)(base + f.out_index16); m.index_type = 1; }
        for (int a = 0; a < f.nattr; a++) {
            const CortoAttr &at = f.attr[a]; uint8_t *p = base + at.out;
            if (at.kind == CK_POSITION) { m.position = (float *)p; bytes_out += (uint64_t)f.nvert * 12; }
            else if (at.kind == CK_UV) { m.uv = (float *)p; bytes_out += (uint64_t)f.nvert * 8; }
            else if (at.kind == CK_NORMAL) { m.normal = (float *)p; bytes_out += (uint64_t)f.nvert * 12; }
            else { m.color = p; bytes_out += (uint64_t)f.nvert * 4; }
        }
    }
    uvol_stats &sx = ctx->stats;
    sx.host_parse_ms = t_parsed - t0; sx.total_ms = now_ms() - t0; sx.kernel_launches = launches; sx.bytes_in = bytes_in; sx.bytes_out = bytes_out; sx.scratch_bytes = zbase + z;
    if (ctx->profile) {
        sx.num_stages = (uint32_t)(ev - 1);
        for (int k = 0; k + 1 < ev && k < 24; k++) cudaEventElapsedTime(&sx.stage_ms[k], ctx->ev[k], ctx->ev[k + 1]);
        float tot = 0; cudaEventElapsedTime(&tot, ctx->ev[1], ctx->ev[ev - 2]); sx.device_ms = tot;
    }
    return UVOL_OK;
}

// ---- the reference's C ABI (corto_codec.h:41-43): one handle = one frame, decoded on device 0 through a
// process-wide context.  Errors are negative return values (the reference lets C++ exceptions escape).
// Normals are written when the file carries them (Decoder::setNormals(float *), corto_codec.cpp:41-44).  Colours: the reference hands
// its `Color *` to the FLOAT path of ColorAttr::dequantize, which multiplies the *unconverted bytes of the buffer* by the channel step
// (color_attribute.cpp:93-106 reads c[k], not the RGB it has just computed) -- an upstream defect whose output is meaningless;
// this entry point returns what that path evidently intends and what the UINT8 path and the TypeScript decoder compute
// (src/lib/corto.ts:455-466): RGBA = toRGB(YCC) * step, as floats in [0, 1].
struct Decoder { std::vector<uint8_t> bytes; uint32_t nvert = 0, nface = 0; };
static uvol_ctx *g_corto_ctx = nullptr;

extern "C" Decoder *CreateDecoder(int length, unsigned char *data, Vector2 *decoderInfo) {
    if (length <= 0 || !data) return nullptr;
    CortoFrame f; memset(&f, 0, sizeof f); std::vector<uint32_t> aux;
    if (corto_parse(data, (size_t)length, f, aux, 1ull << 27) != UVOL_OK && !(f.nvert && f.nface == 0)) return nullptr;
    Decoder *d = new Decoder(); d->bytes.assign(data, data + length); d->nvert = f.nvert; d->nface = f.nface;
    if (decoderInfo) { decoderInfo[0].x = (float)f.nface; decoderInfo[0].y = (float)f.nvert; }
    return d;
}
extern "C" void DestroyDecoder(Decoder *decoder) { delete decoder; }
extern "C" int DecodeMesh(Decoder *decoder, Vector3 *vertices, int *indices, Vector3 *normals, Color *colors, Vector2 *texcoord) {
    if (!decoder) return UVOL_ERR_ARG;
    if (decoder->nface == 0) return -1;                         // point clouds (corto_codec.cpp:27-30)
    if (!g_corto_ctx && uvol_create(0, &g_corto_ctx) != UVOL_OK) return UVOL_ERR_CUDA;
    const uint8_t *p = decoder->bytes.data(); const size_t sz = decoder->bytes.size(); uvol_corto_mesh m;
    const int rc = uvol_decode_corto_batch(g_corto_ctx, &p, &sz, 1, UVOL_MEM_HOST, &m);
    if (rc) return rc;
    if (m.status) return m.status;
    if (indices) memcpy(indices, m.index, (size_t)m.num_faces * 12);
    if (vertices && m.position) memcpy(vertices, m.position, (size_t)m.num_vertices * 12);
    if (texcoord && m.uv) memcpy(texcoord, m.uv, (size_t)m.num_vertices * 8);
    if (normals && m.normal) memcpy(normals, m.normal, (size_t)m.num_vertices * 12);
    if (colors && m.color) for (uint32_t i = 0; i < m.num_vertices; i++) { colors[i].r = m.color[4 * i] / 255.0f; colors[i].g = m.color[4 * i + 1] / 255.0f; colors[i].b = m.color[4 * i + 2] / 255.0f; colors[i].a = m.color[4 * i + 3] / 255.0f; }
    return (int)m.num_faces;
}
