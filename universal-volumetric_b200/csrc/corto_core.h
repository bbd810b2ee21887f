// corto_core.h -- per-unit logic of the V1 (Corto) decode shared by the sm_100a kernels (corto_decode.cu) and the host emulation
// (tests/tools/corto_emu.cpp: logic checks without a GPU; the library has no host decode).
//
// The connectivity walk below produces exactly what crt::Decoder::decodeFaces produces (deprecated/encoder/dev/src/decoder.cpp:181-333
// == src/lib/corto.ts:142-297: the face list and, per new vertex, the parallelogram context), but it is organised for a GPU lane
// rather than around std::vectors of 32-byte edges:
//   * the front is a doubly linked ring of 16-byte records {v0, v1, prev, next}; the third vertex of an edge's face (needed only
//     when a VERTEX symbol creates a vertex) and the "deleted" mark live in a separate word;
//   * the GATE edge and the records of its two ring neighbours stay in registers from one symbol to the next.  The walk is
//     depth first -- the edge a symbol creates is the next gate -- so after a VERTEX symbol (half of all symbols) nothing is
//     loaded at all: the new gate and its right neighbour were just built, the left neighbour is the old one; LEFT / RIGHT
//     need one neighbour of a neighbour, loaded when (and only if) a later symbol asks for it;
//   * the most recent front records are also kept in a small RING (shared memory in the kernel): the walk is depth first, so the
//     neighbour records it asks for are almost always among the last few hundred created -- a shared-memory read instead of a trip
//     to L2 for a line this same lane has just written (ncu, before the ring: 35 % of the walk's time was the wait for front[E.prev]);
//     the array in global memory stays complete (write-through) and serves the old edges the queue brings back;
//   * link updates patch the global record, the ring copy if there is one, and whichever register copy holds that edge;
//   * CLERS symbols are taken eight at a time from an aligned 64-bit word, the next word requested while this one is in use.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CORTO_HD __host__ __device__ __forceinline__
#else
#define CORTO_HD static inline
#endif

enum { CL_VERTEX = 0, CL_LEFT = 1, CL_RIGHT = 2, CL_END = 3, CL_BOUNDARY = 4, CL_DELAY = 5, CL_SPLIT = 6 };
enum { CORTO_OK = 0, CORTO_TRUNCATED = -1, CORTO_CORRUPT = -2 };

struct CortoEdge { int v0, v1, prev, next; };          // 16 bytes: one front edge (v0 -> v1) and its ring links
#define CORTO_DELETED 0x80000000u

// MSB-first bit reader over 32-bit little-endian words (bitstream.cpp:103-121); one word past the stream is readable (padded blob)
struct CortoBits { const uint32_t *w; uint64_t pos, end; };
CORTO_HD bool corto_bits(CortoBits &b, int n, uint32_t *out) {
    if (n == 0) { *out = 0; return true; }
    if (b.pos + (uint64_t)n > b.end) return false;
    const uint64_t wi = b.pos >> 5; const int sh = (int)(b.pos & 31);
    const uint64_t two = ((uint64_t)b.w[wi] << 32) | b.w[wi + 1];
    b.pos += (uint64_t)n;
    *out = (uint32_t)((two << sh) >> (64 - n));
    return true;
}
CORTO_HD int corto_ilog2(uint32_t p) { int k = 0; while (p >>= 1) ++k; return k; }

struct CortoWalkMem {
    const uint8_t *clers; uint32_t nclers;              // symbols; 8-byte aligned and readable up to the next multiple of 8
    CortoBits bits;                                     // split vertices
    const uint32_t *group_end; uint32_t ngroups;        // end face of every group
    CortoEdge *front; uint32_t *third; int front_cap;   // front records; third vertex of the edge's face, or CORTO_DELETED
    CortoEdge *ring; int ring_size;                     // copies of the last ring_size front records (record i at i & (ring_size - 1)); power of two
    int *queue, *delayed; int order_cap;                // edges left behind by VERTEX symbols (first in, first out) / the DELAY stack
    uint32_t *faces;                                    // out: 3 * nface vertex ids
    int *pred;                                          // out: {a, b, c, 0} per vertex -- the parallelogram context of deltaDecode
    int nvert, nface;
    volatile int *progress;                             // optional: number of vertices whose context is final, published every 16 vertices (the kernel's delta warps follow it)
};

CORTO_HD int corto_walk(const CortoWalkMem &m) {
    const int nvert = m.nvert, splitbits = corto_ilog2((uint32_t)nvert) + 1;
    CortoBits bits = m.bits;
    CortoEdge *front = m.front; uint32_t *third = m.third;
    int vertex_count = 0; uint32_t cler = 0, start = 0;
    CortoEdge *ring = m.ring; const int rsize = m.ring_size, rmask = m.ring_size - 1;
    unsigned long long cw = 0, cwn = m.nclers ? *(const unsigned long long *)m.clers : 0ull;
#if defined(__CUDA_ARCH__)
#define CW_PUBLISH() do { if (m.progress && (vertex_count & 15) == 0) { __threadfence_block(); *m.progress = vertex_count; } } while (0)
#else
#define CW_PUBLISH() do { if (m.progress && (vertex_count & 15) == 0) *m.progress = vertex_count; } while (0)
#endif
#define CW_FAIL(code) return (code)
#define CW_SYMBOL(c) do { if (cler >= m.nclers) CW_FAIL(CORTO_TRUNCATED); \
                          if ((cler & 7u) == 0u) { cw = cwn; if (cler + 8u < m.nclers) cwn = *(const unsigned long long *)(m.clers + cler + 8u); } \
                          (c) = (int)((cw >> (8u * (cler & 7u))) & 255u); cler++; } while (0)
// record i lives in the ring as long as fewer than ring_size records were created after it (nfront counts the records created,
// including the ones whose slots are about to be written)
#define CW_IN_RING(i) ((i) >= nfront - rsize)
// (In the kernel the ring is addressed as shared memory outright: through the generic pointer every access paid for an address-space
// conversion -- S2R / R2UR / ULEA on the uniform datapath -- on a chain where each instruction's latency counts.)
#if defined(__CUDA_ARCH__)
    uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    asm volatile("" : "+r"(ring_s));      // opaque: keeps the address in a register (the compiler would otherwise re-derive it, S2UR + ULEA, at every access)
#define CW_RING_GET(dst, i_) asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"((dst).v0), "=r"((dst).v1), "=r"((dst).prev), "=r"((dst).next) : "r"(ring_s + 16u * (uint32_t)((i_) & rmask)))
#define CW_RING_PUT(i_, e) asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" :: "r"(ring_s + 16u * (uint32_t)((i_) & rmask)), "r"((e).v0), "r"((e).v1), "r"((e).prev), "r"((e).next) : "memory")
#define CW_RING_PUT_NEXT(i_, y) asm volatile("st.shared.s32 [%0], %1;" :: "r"(ring_s + 16u * (uint32_t)((i_) & rmask) + 12u), "r"(y) : "memory")
#define CW_RING_PUT_PREV(i_, y) asm volatile("st.shared.s32 [%0], %1;" :: "r"(ring_s + 16u * (uint32_t)((i_) & rmask) + 8u), "r"(y) : "memory")
#else
#define CW_RING_GET(dst, i_) (dst) = ring[(i_) & rmask]
#define CW_RING_PUT(i_, e) ring[(i_) & rmask] = (e)
#define CW_RING_PUT_NEXT(i_, y) ring[(i_) & rmask].next = (y)
#define CW_RING_PUT_PREV(i_, y) ring[(i_) & rmask].prev = (y)
#endif
#define CW_LOAD(dst, i) do { const int i_ = (i); if (CW_IN_RING(i_)) CW_RING_GET(dst, i_); else (dst) = front[i_]; } while (0)
#define CW_NEW(i, e) do { const int i_ = (i); CW_RING_PUT(i_, e); front[i_] = (e); } while (0)
#define CW_SET_NEXT(x, y) do { const int x_ = (x); front[x_].next = (y); if (CW_IN_RING(x_)) CW_RING_PUT_NEXT(x_, y); if (x_ == pi) P.next = (y); if (x_ == ni) N.next = (y); } while (0)
#define CW_SET_PREV(x, y) do { const int x_ = (x); front[x_].prev = (y); if (CW_IN_RING(x_)) CW_RING_PUT_PREV(x_, y); if (x_ == pi) P.prev = (y); if (x_ == ni) N.prev = (y); } while (0)
#define CW_FACE(a, b, c) do { if (start + 3 > end) CW_FAIL(CORTO_CORRUPT); m.faces[start] = (uint32_t)(a); m.faces[start + 1] = (uint32_t)(b); m.faces[start + 2] = (uint32_t)(c); start += 3; } while (0)
    for (uint32_t gi = 0; gi < m.ngroups; gi++) {
        const uint32_t end = m.group_end[gi] * 3u;
        int nfront = 0, norder = 0, order = 0, ndelayed = 0;
        int g = -1, pi = -1, ni = -1; uint32_t gv2 = 0;          // gate edge, which edges the neighbour copies hold, the gate's third vertex
        CortoEdge E = {0, 0, 0, 0}, P = {0, 0, 0, 0}, N = {0, 0, 0, 0};
        while (start < end) {
            if (g < 0) {
                if (order < norder) g = m.queue[order++];
                else if (ndelayed) g = m.delayed[--ndelayed];
                else {
                    // ---- a new component: one face, three front edges
                    int c; CW_SYMBOL(c);
                    uint32_t split = 0;
                    if (c == CL_SPLIT) { if (!corto_bits(bits, 3, &split)) CW_FAIL(CORTO_TRUNCATED); }
                    else if (c != CL_VERTEX) CW_FAIL(CORTO_CORRUPT);
                    int last = vertex_count - 1, v[3];
                    for (int k = 0; k < 3; k++) {
                        if (split & (1u << k)) { uint32_t s; if (!corto_bits(bits, splitbits, &s)) CW_FAIL(CORTO_TRUNCATED); if ((int)s >= nvert) CW_FAIL(CORTO_CORRUPT); v[k] = (int)s; }
                        else {
                            if (vertex_count >= nvert) CW_FAIL(CORTO_CORRUPT);
                            int *pr = m.pred + 4 * vertex_count; pr[0] = last; pr[1] = last; pr[2] = last; pr[3] = 0;
                            last = v[k] = vertex_count++;
                            CW_PUBLISH();
                        }
                    }
                    CW_FACE(v[0], v[1], v[2]);
                    const int cur = nfront;
                    if (nfront + 3 > m.front_cap || norder + 3 > m.order_cap) CW_FAIL(CORTO_CORRUPT);
                    nfront += 3;
                    { const CortoEdge e0 = {v[1], v[2], cur + 2, cur + 1}, e1 = {v[2], v[0], cur, cur + 2}, e2 = {v[0], v[1], cur + 1, cur};
                      CW_NEW(cur, e0); CW_NEW(cur + 1, e1); CW_NEW(cur + 2, e2); }
                    third[cur] = (uint32_t)v[0]; third[cur + 1] = (uint32_t)v[1]; third[cur + 2] = (uint32_t)v[2];
                    m.queue[norder++] = cur; m.queue[norder++] = cur + 1; m.queue[norder++] = cur + 2;
                    continue;
                }
                gv2 = third[g];
                if (gv2 & CORTO_DELETED) { g = -1; continue; }
                CW_LOAD(E, g); pi = ni = -1;
            }
            int c; CW_SYMBOL(c);
            if (c == CL_VERTEX || c == CL_SPLIT) {
                int opp;
                if (c == CL_SPLIT) { uint32_t s; if (!corto_bits(bits, splitbits, &s)) CW_FAIL(CORTO_TRUNCATED); opp = (int)s; if (opp >= nvert) CW_FAIL(CORTO_CORRUPT); }
                else {
                    if (vertex_count >= nvert) CW_FAIL(CORTO_CORRUPT);
                    int *pr = m.pred + 4 * vertex_count; pr[0] = E.v1; pr[1] = E.v0; pr[2] = (int)gv2; pr[3] = 0;
                    opp = vertex_count++;
                    CW_PUBLISH();
                }
                if (nfront + 2 > m.front_cap || norder >= m.order_cap) CW_FAIL(CORTO_CORRUPT);
                const int A = nfront, B = nfront + 1; nfront += 2;
                CW_SET_NEXT(E.prev, A); CW_SET_PREV(E.next, B);
                const CortoEdge ea = {E.v0, opp, E.prev, B}, eb = {opp, E.v1, A, E.next};
                CW_NEW(A, ea); third[A] = (uint32_t)E.v1; CW_NEW(B, eb); third[B] = (uint32_t)E.v0;
                m.queue[norder++] = B;
                CW_FACE(E.v1, E.v0, opp);
                // the left edge of the new face is the next gate: its left neighbour is the old one (kept if it was held), its right
                // neighbour is the edge just built
                if (pi != E.prev) pi = -1;
                gv2 = (uint32_t)E.v1; g = A; N = eb; ni = B; E = ea;
            } else if (c == CL_LEFT) {
                if (pi != E.prev) { CW_LOAD(P, E.prev); pi = E.prev; }
                third[E.prev] = CORTO_DELETED;
                if (nfront + 1 > m.front_cap) CW_FAIL(CORTO_CORRUPT);
                const int X = nfront++, opp = P.v0, pp = P.prev;
                CW_SET_NEXT(pp, X); CW_SET_PREV(E.next, X);
                const CortoEdge ex = {opp, E.v1, pp, E.next};
                CW_NEW(X, ex); third[X] = (uint32_t)E.v0;
                CW_FACE(E.v1, E.v0, opp);
                if (ni != E.next) ni = -1;
                if (pp == ni) { P = N; pi = ni; } else pi = -1;
                gv2 = (uint32_t)E.v0; g = X; E = ex;
            } else if (c == CL_RIGHT) {
                if (ni != E.next) { CW_LOAD(N, E.next); ni = E.next; }
                third[E.next] = CORTO_DELETED;
                if (nfront + 1 > m.front_cap) CW_FAIL(CORTO_CORRUPT);
                const int X = nfront++, opp = N.v1, nn = N.next;
                CW_SET_PREV(nn, X); CW_SET_NEXT(E.prev, X);
                const CortoEdge ex = {E.v0, opp, E.prev, nn};
                CW_NEW(X, ex); third[X] = (uint32_t)E.v1;
                CW_FACE(E.v1, E.v0, opp);
                if (pi != E.prev) pi = -1;
                if (nn == pi) { N = P; ni = pi; } else ni = -1;
                gv2 = (uint32_t)E.v1; g = X; E = ex;
            } else if (c == CL_END) {
                if (pi != E.prev) { CW_LOAD(P, E.prev); pi = E.prev; }
                if (ni != E.next) { CW_LOAD(N, E.next); ni = E.next; }
                third[E.prev] = CORTO_DELETED; third[E.next] = CORTO_DELETED;
                const int pp = P.prev, nn = N.next, opp = P.v0;
                pi = ni = -1;
                CW_SET_NEXT(pp, nn); CW_SET_PREV(nn, pp);
                CW_FACE(E.v1, E.v0, opp);
                g = -1;
            } else if (c == CL_BOUNDARY) g = -1;
            else if (c == CL_DELAY) { if (ndelayed >= m.order_cap) CW_FAIL(CORTO_CORRUPT); m.delayed[ndelayed++] = g; g = -1; }
            else CW_FAIL(CORTO_CORRUPT);
        }
    }
#undef CW_PUBLISH
#undef CW_FAIL
#undef CW_SYMBOL
#undef CW_IN_RING
#undef CW_RING_GET
#undef CW_RING_PUT
#undef CW_RING_PUT_NEXT
#undef CW_RING_PUT_PREV
#undef CW_LOAD
#undef CW_NEW
#undef CW_SET_NEXT
#undef CW_SET_PREV
#undef CW_FACE
    return vertex_count == nvert ? CORTO_OK : CORTO_CORRUPT;
}
