// corto_decode.cu -- sm_100a kernels + launcher for the V1 geometry path (Corto .crt frames), and the
// reference's own C ABI (CreateDecoder / DecodeMesh / DestroyDecoder, deprecated/encoder/dev/src/corto_codec.h:41-43).
//
// Replaces CortoDecoder.decode (src/lib/corto.ts:73-140) == crt::Decoder::decode
// (deprecated/encoder/dev/src/decoder.cpp:122-173) for a batch of frames sliced out of a .drcs
// (src/V1/worker.ts:48-68).  Layout: SURVEY.md Appendix C.  Stages:
//   tunstall   one warp per Tunstall block: lane 0 rebuilds the 256-word dictionary (tunstall.cpp:125-256),
//              then the warp expands the code bytes with a prefix sum over word lengths (:430-452)
//   faces      one warp per frame: the front-growing connectivity walk (decoder.cpp:181-333), serial
//   values     one CTA per (frame, attribute): prefix sum of the per-value bit widths -> every value's bit
//              offset -> parallel bit extraction (cstream.h:296-362; bitstream.cpp:103-121, MSB-first words)
//   delta      one warp per (frame, attribute): parallelogram / delta reversal (vertex_attribute.h:155-177)
//   dequant    element-parallel: (float)value * q (vertex_attribute.h:179-224)
// Supported: generic attributes "position" (3 x f32) and "uv" (2 x f32), both strategies, Tunstall or no
// entropy coding; normal / colour codecs are rejected with UNSUPPORTED (UVOL V1 carries position + uv only,
// src/V1/player.ts:292-294).
#include <chrono>
#include <string.h>
#include <string>
#include <vector>
#include "uvol_ctx.h"
#include "../../include/corto_codec.h"

namespace {

enum { CL_VERTEX = 0, CL_LEFT = 1, CL_RIGHT = 2, CL_END = 3, CL_BOUNDARY = 4, CL_DELAY = 5, CL_SPLIT = 6 };

struct TunBlock { uint32_t probs_off, nsym, size, csize, data_off; uint64_t o_out; };      // offsets inside the file; o_out in scratch
struct BitBlock { uint32_t nwords, data_off; };
struct CortoAttr { int32_t kind /*0 position 1 uv*/, N, strategy, nlogs; float q; BitBlock bits; TunBlock logs[4]; uint64_t o_val, out; };
struct CortoFrame {
    uint64_t file_off; uint32_t file_len; int32_t status;
    uint32_t nvert, nface, ngroups, groups_off /*aux u32: end face per group*/, max_front, entropy;
    TunBlock clers; BitBlock ibits;
    int32_t nattr; CortoAttr attr[2];
    uint64_t o_front, o_order, o_delayed, o_pred, out_index;
};
struct CJob { uint32_t frame; int32_t what; };

struct Rd {
    const uint8_t *b; size_t n, p; bool err;
    uint8_t u8() { if (p + 1 > n) { err = true; return 0; } return b[p++]; }
    uint16_t u16() { uint16_t a = u8(), c = u8(); return (uint16_t)(a | (c << 8)); }
    uint32_t u32() { if (p + 4 > n) { err = true; p = n; return 0; } uint32_t v; memcpy(&v, b + p, 4); p += 4; return v; }
    float f32() { uint32_t v = u32(); float f; memcpy(&f, &v, 4); return f; }
    std::string str() { uint16_t l = u16(); if (err || p + l > n) { err = true; return ""; } std::string s((const char *)b + p, l ? l - 1 : 0); p += l; return s; }
};
bool read_tunstall(Rd &r, uint32_t entropy, TunBlock &t) {
    memset(&t, 0, sizeof t);
    if (entropy == 0) { t.nsym = 0xffffffffu; t.size = r.u32(); t.csize = t.size; t.data_off = (uint32_t)r.p; if (r.err || r.p + t.size > r.n) return false; r.p += t.size; return true; }
    t.nsym = r.u8(); t.probs_off = (uint32_t)r.p; r.p += 2 * (size_t)t.nsym;
    t.size = r.u32(); t.csize = r.u32(); t.data_off = (uint32_t)r.p;
    if (r.err || r.p + t.csize > r.n || t.size > (1u << 28)) return false;
    r.p += t.csize;
    return true;
}
bool read_bits(Rd &r, BitBlock &b) {
    b.nwords = r.u32();
    const size_t pad = r.p & 3; if (pad) r.p += 4 - pad;
    b.data_off = (uint32_t)r.p;
    if (r.err || r.p + 4ull * b.nwords > r.n) return false;
    r.p += 4ull * b.nwords;
    return true;
}

// Header + section walk (decoder.cpp:41-85, index_attribute.h:83-98, cstream.h:285-362).
int corto_parse(const uint8_t *data, size_t len, CortoFrame &f, std::vector<uint32_t> &aux) {
    Rd r{data, len, 0, false};
    if (len < 24 || r.u32() != 0x787A6300u) return UVOL_ERR_CORRUPT;
    (void)r.u32();
    f.entropy = r.u8();
    if (f.entropy > 1) return UVOL_ERR_UNSUPPORTED;
    const uint32_t nexif = r.u32();
    if (r.err || nexif > 4096) return UVOL_ERR_CORRUPT;
    for (uint32_t i = 0; i < nexif; i++) { r.str(); r.str(); }
    const uint32_t nattr = r.u32();
    if (r.err || nattr > 16) return UVOL_ERR_CORRUPT;
    struct Hdr { std::string name; int codec; float q; int N, format, strategy; };
    std::vector<Hdr> hdr(nattr);
    for (auto &h : hdr) { h.name = r.str(); h.codec = (int)r.u32(); h.q = r.f32(); h.N = r.u8(); h.format = r.u8(); h.strategy = r.u8(); }
    f.nvert = r.u32(); f.nface = r.u32();
    if (r.err || f.nvert == 0 || f.nvert > (1u << 26) || f.nface > (1u << 27)) return UVOL_ERR_CORRUPT;
    if (f.nface == 0) return UVOL_ERR_UNSUPPORTED;                 // point clouds: DecodeMesh returns -1 (corto_codec.cpp:27-30)
    f.ngroups = r.u32(); f.groups_off = (uint32_t)aux.size();
    if (r.err || f.ngroups > 65536) return UVOL_ERR_CORRUPT;
    for (uint32_t g = 0; g < f.ngroups; g++) {
        const uint32_t end = r.u32(); const uint8_t np = r.u8();
        for (int k = 0; k < np; k++) { r.str(); r.str(); }
        if (r.err || end > f.nface) return UVOL_ERR_CORRUPT;
        aux.push_back(end);
    }
    f.max_front = r.u32();
    if (!read_tunstall(r, f.entropy, f.clers) || !read_bits(r, f.ibits)) return UVOL_ERR_TRUNCATED;
    // attributes follow in std::map (alphabetical) order of their names (decoder.cpp:146-147)
    std::vector<int> order(nattr); for (uint32_t i = 0; i < nattr; i++) order[i] = (int)i;
    for (uint32_t i = 0; i < nattr; i++) for (uint32_t j = i + 1; j < nattr; j++) if (hdr[order[j]].name < hdr[order[i]].name) std::swap(order[i], order[j]);
    f.nattr = 0;
    for (uint32_t i = 0; i < nattr; i++) {
        const Hdr &h = hdr[order[i]];
        if (h.codec != 1 || h.N < 1 || h.N > 4) return UVOL_ERR_UNSUPPORTED;
        CortoAttr a; memset(&a, 0, sizeof a);
        a.kind = h.name == "position" ? 0 : (h.name == "uv" ? 1 : -1); a.N = h.N; a.strategy = h.strategy; a.q = h.q;
        if (!read_bits(r, a.bits)) return UVOL_ERR_TRUNCATED;
        a.nlogs = (h.strategy & 2) ? 1 : h.N;
        for (int k = 0; k < a.nlogs; k++) { if (!read_tunstall(r, f.entropy, a.logs[k])) return UVOL_ERR_TRUNCATED; if (a.logs[k].size != f.nvert) return UVOL_ERR_CORRUPT; }
        if (a.kind < 0 || (a.kind == 0 && a.N != 3) || (a.kind == 1 && a.N != 2) || f.nattr >= 2) return UVOL_ERR_UNSUPPORTED;
        f.attr[f.nattr++] = a;
    }
    if (f.nattr < 1 || f.attr[0].kind != 0) return UVOL_ERR_UNSUPPORTED;
    return UVOL_OK;
}

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

// MSB-first bit reader over 32-bit little-endian words (bitstream.cpp:103-121)
struct BitsMsb { const uint32_t *w; uint64_t pos; };
__device__ __forceinline__ uint32_t bits_read(BitsMsb &b, int n) {
    if (n == 0) return 0;
    const uint64_t wi = b.pos >> 5; const int sh = (int)(b.pos & 31);
    const uint64_t two = ((uint64_t)b.w[wi] << 32) | b.w[wi + 1];
    b.pos += (uint64_t)n;
    return (uint32_t)((two << sh) >> (64 - n));
}
__device__ __forceinline__ int ilog2_u(uint32_t p) { int k = 0; while (p >>= 1) ++k; return k; }

struct FrontEdge { int v0, v1, v2, prev, next, deleted, pad0, pad1; };

// Connectivity: front-growing walk (decoder.cpp:181-333), one warp per frame, lane 0 walks.
__global__ void __launch_bounds__(32 * CW) k_corto_faces(const CortoFrame *frames, int32_t *status, const uint8_t *blob, const uint32_t *aux, uint8_t *S, uint8_t *O, int nframes) {
    const int fi = blockIdx.x * CW + (threadIdx.x >> 5);
    if (fi >= nframes || (threadIdx.x & 31) != 0) return;
    if (frames[fi].status) { status[fi] = frames[fi].status; return; }
    if (status[fi]) return;
    const CortoFrame &f = frames[fi];
    const uint8_t *clers = S + f.clers.o_out; const uint32_t nclers = f.clers.size;
    BitsMsb bits{(const uint32_t *)(blob + f.file_off + f.ibits.data_off), 0};
    const uint64_t maxbits = (uint64_t)f.ibits.nwords * 32;
    FrontEdge *front = (FrontEdge *)(S + f.o_front); int *faceorder = (int *)(S + f.o_order), *delayed = (int *)(S + f.o_delayed);
    int4 *pred = (int4 *)(S + f.o_pred); uint32_t *faces = (uint32_t *)(O + f.out_index);
    const int nvert = (int)f.nvert, splitbits = ilog2_u(f.nvert) + 1;
    const int front_cap = 3 * (int)f.nface + 8, order_cap = 2 * (int)f.nface + 8;
    int vertex_count = 0, st = 0; uint32_t cler = 0, start = 0;
#define CFAIL(code) do { st = (code); goto done; } while (0)
#define PUSH_EDGE(a, b, c, p, n) do { if (nfront >= front_cap) CFAIL(UVOL_ERR_CORRUPT); FrontEdge e_ = {(a), (b), (c), (p), (n), 0, 0, 0}; front[nfront++] = e_; } while (0)
    for (uint32_t g = 0; g < f.ngroups; g++) {
        const uint32_t end = aux[f.groups_off + g] * 3;
        int nfront = 0, norder = 0, order = 0, ndelayed = 0, new_edge = -1;
        while (start < end) {
            if (new_edge == -1 && order >= norder && ndelayed == 0) {
                int last_index = vertex_count - 1, vindex[3], split = 0;
                if (cler >= nclers) CFAIL(UVOL_ERR_TRUNCATED);
                const int c = clers[cler++];
                if (c == CL_SPLIT) { if (bits.pos + 3 > maxbits) CFAIL(UVOL_ERR_TRUNCATED); split = (int)bits_read(bits, 3); }
                else if (c != CL_VERTEX) CFAIL(UVOL_ERR_CORRUPT);
                for (int k = 0; k < 3; k++) {
                    int v;
                    if (split & (1 << k)) { if (bits.pos + splitbits > maxbits) CFAIL(UVOL_ERR_TRUNCATED); v = (int)bits_read(bits, splitbits); if (v >= nvert) CFAIL(UVOL_ERR_CORRUPT); }
                    else { if (vertex_count >= nvert) CFAIL(UVOL_ERR_CORRUPT); pred[vertex_count] = make_int4(last_index, last_index, last_index, 0); last_index = v = vertex_count++; }
                    vindex[k] = v; faces[start++] = (uint32_t)v;
                }
                const int cur = nfront;
                if (norder + 3 > order_cap) CFAIL(UVOL_ERR_CORRUPT);
                faceorder[norder++] = nfront; PUSH_EDGE(vindex[1], vindex[2], vindex[0], cur + 2, cur + 1);
                faceorder[norder++] = nfront; PUSH_EDGE(vindex[2], vindex[0], vindex[1], cur + 0, cur + 2);
                faceorder[norder++] = nfront; PUSH_EDGE(vindex[0], vindex[1], vindex[2], cur + 1, cur + 0);
                continue;
            }
            int fe;
            if (new_edge != -1) { fe = new_edge; new_edge = -1; }
            else if (order < norder) fe = faceorder[order++];
            else if (ndelayed) fe = delayed[--ndelayed];
            else CFAIL(UVOL_ERR_CORRUPT);
            const FrontEdge e = front[fe];
            if (e.deleted) continue;
            if (cler >= nclers) CFAIL(UVOL_ERR_TRUNCATED);
            const int c = clers[cler++];
            if (c == CL_BOUNDARY) continue;
            const int v0 = e.v0, v1 = e.v1;
            const FrontEdge pe = front[e.prev], ne = front[e.next];
            new_edge = nfront; int opposite = -1;
            if (c == CL_VERTEX || c == CL_SPLIT) {
                if (c == CL_SPLIT) { if (bits.pos + splitbits > maxbits) CFAIL(UVOL_ERR_TRUNCATED); opposite = (int)bits_read(bits, splitbits); }
                else { if (vertex_count >= nvert) CFAIL(UVOL_ERR_CORRUPT); pred[vertex_count] = make_int4(v1, v0, e.v2, 0); opposite = vertex_count++; }
                if (opposite >= nvert) CFAIL(UVOL_ERR_CORRUPT);
                front[e.prev].next = new_edge; front[e.next].prev = new_edge + 1;
                PUSH_EDGE(v0, opposite, v1, e.prev, new_edge + 1);
                if (norder >= order_cap) CFAIL(UVOL_ERR_CORRUPT);
                faceorder[norder++] = nfront;
                PUSH_EDGE(opposite, v1, v0, new_edge, e.next);
            } else if (c == CL_LEFT) {
                front[e.prev].deleted = 1; front[pe.prev].next = new_edge; front[e.next].prev = new_edge; opposite = pe.v0;
                PUSH_EDGE(opposite, v1, v0, pe.prev, e.next);
            } else if (c == CL_RIGHT) {
                front[e.next].deleted = 1; front[ne.next].prev = new_edge; front[e.prev].next = new_edge; opposite = ne.v1;
                PUSH_EDGE(v0, opposite, v1, e.prev, ne.next);
            } else if (c == CL_DELAY) {
                if (ndelayed >= order_cap) CFAIL(UVOL_ERR_CORRUPT);
                delayed[ndelayed++] = fe; new_edge = -1; continue;
            } else if (c == CL_END) {
                front[e.prev].deleted = 1; front[e.next].deleted = 1; front[pe.prev].next = ne.next; front[ne.next].prev = pe.prev; opposite = pe.v0; new_edge = -1;
            } else CFAIL(UVOL_ERR_CORRUPT);
            if (start + 3 > end) CFAIL(UVOL_ERR_CORRUPT);
            faces[start++] = (uint32_t)v1; faces[start++] = (uint32_t)v0; faces[start++] = (uint32_t)opposite;
        }
    }
    if (vertex_count != nvert) st = UVOL_ERR_CORRUPT;
done:
    if (st) status[fi] = st;
#undef CFAIL
#undef PUSH_EDGE
}

// Per-value bit widths -> bit offsets -> values (cstream.h:296-362).  grid = (frames, attrs), 256 threads.
__global__ void __launch_bounds__(256) k_corto_values(const CortoFrame *frames, const int32_t *status, const uint8_t *blob, uint8_t *S) {
    __shared__ unsigned long long wsum[8], carry_s;
    const uint32_t fi = blockIdx.x, ai = blockIdx.y;
    if (frames[fi].status || status[fi]) return;
    const CortoFrame &f = frames[fi];
    if ((int)ai >= f.nattr) return;
    const CortoAttr &a = f.attr[ai];
    const uint32_t *words = (const uint32_t *)(blob + f.file_off + a.bits.data_off);
    int32_t *val = (int32_t *)(S + a.o_val);
    const int n = (int)f.nvert, N = a.N, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool corr = (a.strategy & 2) != 0;
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
// or += values[a].  One warp per (frame, attribute), lane k owns component k.
__global__ void __launch_bounds__(32 * CW) k_corto_delta(const CortoFrame *frames, const int32_t *status, uint8_t *S, const CJob *jobs, int njobs) {
    const int ji = blockIdx.x * CW + (threadIdx.x >> 5), k = threadIdx.x & 31;
    if (ji >= njobs) return;
    const CJob jb = jobs[ji];
    if (frames[jb.frame].status || status[jb.frame]) return;
    const CortoFrame &f = frames[jb.frame]; const CortoAttr &a = f.attr[jb.what];
    if (k >= a.N) return;
    int32_t *val = (int32_t *)(S + a.o_val); const int4 *pred = (const int4 *)(S + f.o_pred);
    const int n = (int)f.nvert, N = a.N; const bool par = (a.strategy & 1) != 0;
    int4 p = n > 1 ? pred[1] : make_int4(0, 0, 0, 0);
    for (int i = 1; i < n; i++) {
        const int4 nx = i + 1 < n ? pred[i + 1] : p;       // prefetch the next context
        int v = val[i * N + k];
        if (par) v += val[p.x * N + k] + val[p.y * N + k] - val[p.z * N + k]; else v += val[p.x * N + k];
        val[i * N + k] = v;
        p = nx;
    }
}

// Dequantisation: (float)value * q, one rounding (vertex_attribute.h:186-187).  grid = (ceil(max/256), frames, attrs)
__global__ void __launch_bounds__(256) k_corto_dequant(const CortoFrame *frames, const int32_t *status, const uint8_t *S, uint8_t *O) {
    const uint32_t fi = blockIdx.y, ai = blockIdx.z;
    if (frames[fi].status || status[fi]) return;
    const CortoFrame &f = frames[fi];
    if ((int)ai >= f.nattr) return;
    const CortoAttr &a = f.attr[ai];
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= f.nvert * (uint32_t)a.N) return;
    ((float *)(O + a.out))[i] = __fmul_rn((float)((const int32_t *)(S + a.o_val))[i], a.q);
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
uint64_t take(uint64_t &cur, uint64_t bytes) { uint64_t o = cur; cur = (cur + bytes + 127) / 128 * 128; return o; }

}  // namespace

struct CortoBatch { std::vector<CortoFrame> frames; };
void uvol_corto_batch_free(CortoBatch *b) { delete b; }
static const char *kCortoStages[] = {"h2d", "tunstall", "faces", "values", "delta", "dequant", "d2h"};
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
    uint64_t blob_bytes = 0, s = 0, o = 0; uint32_t maxvals = 1; uint64_t bytes_in = 0;
    for (int i = 0; i < n; i++) {
        CortoFrame &f = frames[i]; memset(&f, 0, sizeof f);
        f.file_off = blob_bytes; f.file_len = (uint32_t)size[i]; bytes_in += size[i];
        blob_bytes = align_up(blob_bytes + size[i] + 16, 16);      // 4-byte alignment of the bit streams is preserved (decoder.cpp:42-43)
        f.status = (data[i] && size[i] < (1ull << 31)) ? corto_parse(data[i], size[i], f, aux) : UVOL_ERR_ARG;
        if (f.status) continue;
        f.clers.o_out = take(s, (uint64_t)f.clers.size + 8);
        f.o_front = take(s, (3ull * f.nface + 8) * sizeof(FrontEdge)); f.o_order = take(s, (2ull * f.nface + 8) * 4); f.o_delayed = take(s, (2ull * f.nface + 8) * 4);
        f.o_pred = take(s, ((uint64_t)f.nvert + 2) * 16);
        f.out_index = take(o, (uint64_t)f.nface * 12);
        for (int a = 0; a < f.nattr; a++) {
            CortoAttr &at = f.attr[a];
            for (int k = 0; k < at.nlogs; k++) at.logs[k].o_out = take(s, (uint64_t)at.logs[k].size + 8);
            at.o_val = take(s, (uint64_t)f.nvert * at.N * 4); at.out = take(o, (uint64_t)f.nvert * at.N * 4);
            if (f.nvert * (uint32_t)at.N > maxvals) maxvals = f.nvert * (uint32_t)at.N;
        }
    }
    aux.push_back(0);
    const int j_tun = 0;
    for (int i = 0; i < n; i++) if (!frames[i].status) { jobs.push_back({(uint32_t)i, 0}); for (int a = 0; a < frames[i].nattr; a++) for (int k = 0; k < frames[i].attr[a].nlogs; k++) jobs.push_back({(uint32_t)i, 1 + 4 * a + k}); }
    const int j_delta = (int)jobs.size();
    for (int i = 0; i < n; i++) if (!frames[i].status) for (int a = 0; a < frames[i].nattr; a++) jobs.push_back({(uint32_t)i, a});
    const int j_end = (int)jobs.size();
    const size_t desc_bytes = sizeof(CortoFrame) * (size_t)n, aux_bytes = aux.size() * 4, job_bytes = sizeof(CJob) * (jobs.size() + 1);
    UVOL_CUDA(ctx, ctx->h_cblob.reserve(blob_bytes + 64)); UVOL_CUDA(ctx, ctx->d_cblob.reserve(blob_bytes + 64));
    UVOL_CUDA(ctx, ctx->h_cdesc.reserve(desc_bytes + aux_bytes + job_bytes + 64)); UVOL_CUDA(ctx, ctx->d_cdesc.reserve(desc_bytes + aux_bytes + job_bytes + 64));
    UVOL_CUDA(ctx, ctx->d_cscratch.reserve(s + 256)); UVOL_CUDA(ctx, ctx->d_out_corto.reserve(o + 256));
    UVOL_CUDA(ctx, ctx->d_ccounts.reserve(4 * (size_t)n)); UVOL_CUDA(ctx, ctx->h_ccounts.reserve(4 * (size_t)n));
    memset(ctx->h_cblob.p, 0, blob_bytes + 64);
    for (int i = 0; i < n; i++) if (data[i] && size[i] < (1ull << 31)) memcpy((uint8_t *)ctx->h_cblob.p + frames[i].file_off, data[i], size[i]);
    uint8_t *hd = (uint8_t *)ctx->h_cdesc.p;
    memcpy(hd, frames.data(), desc_bytes); memcpy(hd + desc_bytes, aux.data(), aux_bytes); memcpy(hd + desc_bytes + aux_bytes, jobs.data(), sizeof(CJob) * jobs.size());
    const double t_parsed = now_ms();
    cudaStream_t st = ctx->s0; int ev = 0;
    auto stamp = [&]() { if (ctx->profile && ev < 32) cudaEventRecord(ctx->ev[ev], st); ev++; };
    stamp();
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_cblob.p, ctx->h_cblob.p, blob_bytes + 64, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemcpyAsync(ctx->d_cdesc.p, hd, desc_bytes + aux_bytes + job_bytes, cudaMemcpyHostToDevice, st));
    UVOL_CUDA(ctx, cudaMemsetAsync(ctx->d_ccounts.p, 0, 4 * (size_t)n, st));
    stamp();
    const CortoFrame *dF = (const CortoFrame *)ctx->d_cdesc.p; const uint32_t *dAux = (const uint32_t *)((const uint8_t *)ctx->d_cdesc.p + desc_bytes);
    const CJob *dJ = (const CJob *)((const uint8_t *)ctx->d_cdesc.p + desc_bytes + aux_bytes);
    int32_t *dSt = (int32_t *)ctx->d_ccounts.p; const uint8_t *dBlob = (const uint8_t *)ctx->d_cblob.p;
    uint8_t *dS = (uint8_t *)ctx->d_cscratch.p, *dO = (uint8_t *)ctx->d_out_corto.p;
    uint32_t launches = 0;
    auto nb = [](int j) { return (unsigned)((j + CW - 1) / CW); };
    if (j_delta - j_tun > 0) {
        UVOL_CUDA(ctx, cudaFuncSetAttribute(k_tunstall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(TunSmem) * CW)));
        k_tunstall<<<nb(j_delta - j_tun), 32 * CW, sizeof(TunSmem) * CW, st>>>(dF, dSt, dBlob, dS, dJ + j_tun, j_delta - j_tun); launches++;
    }
    stamp();
    k_corto_faces<<<nb(n), 32 * CW, 0, st>>>(dF, dSt, dBlob, dAux, dS, dO, n); launches++;
    stamp();
    k_corto_values<<<dim3(n, 2), 256, 0, st>>>(dF, dSt, dBlob, dS); launches++;
    stamp();
    if (j_end - j_delta > 0) { k_corto_delta<<<nb(j_end - j_delta), 32 * CW, 0, st>>>(dF, dSt, dS, dJ + j_delta, j_end - j_delta); launches++; }
    stamp();
    k_corto_dequant<<<dim3((maxvals + 255) / 256, n, 2), 256, 0, st>>>(dF, dSt, dS, dO); launches++;
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
        for (int a = 0; a < f.nattr; a++) { float *p = (float *)(base + f.attr[a].out); bytes_out += (uint64_t)f.nvert * f.attr[a].N * 4; if (f.attr[a].kind == 0) m.position = p; else m.uv = p; }
    }
    uvol_stats &sx = ctx->stats;
    sx.host_parse_ms = t_parsed - t0; sx.total_ms = now_ms() - t0; sx.kernel_launches = launches; sx.bytes_in = bytes_in; sx.bytes_out = bytes_out; sx.scratch_bytes = s;
    if (ctx->profile) {
        sx.num_stages = (uint32_t)(ev - 1);
        for (int k = 0; k + 1 < ev && k < 24; k++) cudaEventElapsedTime(&sx.stage_ms[k], ctx->ev[k], ctx->ev[k + 1]);
        float tot = 0; cudaEventElapsedTime(&tot, ctx->ev[1], ctx->ev[ev - 2]); sx.device_ms = tot;
    }
    return UVOL_OK;
}

// ---- the reference's C ABI (corto_codec.h:41-43): one handle = one frame, decoded on device 0 through a
// process-wide context.  Errors are negative return values (the reference lets C++ exceptions escape).
struct Decoder { std::vector<uint8_t> bytes; uint32_t nvert = 0, nface = 0; bool has_uv = false; };
static uvol_ctx *g_corto_ctx = nullptr;

extern "C" Decoder *CreateDecoder(int length, unsigned char *data, Vector2 *decoderInfo) {
    if (length <= 0 || !data) return nullptr;
    CortoFrame f; memset(&f, 0, sizeof f); std::vector<uint32_t> aux;
    if (corto_parse(data, (size_t)length, f, aux) != UVOL_OK && !(f.nvert && f.nface == 0)) return nullptr;
    Decoder *d = new Decoder(); d->bytes.assign(data, data + length); d->nvert = f.nvert; d->nface = f.nface;
    for (int a = 0; a < f.nattr; a++) if (f.attr[a].kind == 1) d->has_uv = true;
    if (decoderInfo) { decoderInfo[0].x = (float)f.nface; decoderInfo[0].y = (float)f.nvert; }
    return d;
}
extern "C" void DestroyDecoder(Decoder *decoder) { delete decoder; }
extern "C" int DecodeMesh(Decoder *decoder, Vector3 *vertices, int *indices, Vector3 *normals, Color *colors, Vector2 *texcoord) {
    (void)normals; (void)colors;
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
    return (int)m.num_faces;
}
