/*
 * draco_oracle.c -- CPU restatement of the Draco mesh decode that the reference's
 * V2 geometry path runs inside its WASM worker.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under universal-volumetric_b200/ may link,
 * import or call this file; it exists so tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs have something to check the
 * CUDA path against and to time on the host cores.
 *
 * What it restates
 *   reference call site : src/lib/DRACOLoader.js:470-554 (decodeGeometry ->
 *                         decoder.DecodeArrayToMesh :483), :556-569 (decodeIndex,
 *                         GetTrianglesUInt32Array), :571-590 (decodeAttribute,
 *                         GetAttributeDataArrayForAllPoints DT_FLOAT32)
 *   arithmetic          : third-party Google Draco, decoder 1.4.3 (WASM fetched from
 *                         gstatic at run time, src/V2/player.ts:101; npm draco3d locked
 *                         1.5.6 in yarn.lock:566-569).  The source is NOT in
 *                         /root/reference, so this file restates the published
 *                         bitstream 2.2 / edgebreaker algorithm.
 *
 * PARITY UNPINNED against upstream binaries: the reference ships no tests and no
 * decoded golden vectors for this path, and no Draco build exists in this image.
 * The substitute pins are the reference's own 250 .drc fixtures
 * (example/public/liam/output/geometry_draco) checked with the self-consistency
 * oracles of SURVEY.md A.4 (parse-to-EOF, rANS terminal state, context counters hit
 * zero, orientation bits consumed exactly, value ranges) -- see tests/test_oracle_draco.py.
 * Unpinned: exact point-id order, fp32 rounding order of dequantisation and of the
 * octahedral->unit-vector step.  Each lives in one small function below
 * (assign_points, dequant_value, oct_to_unit) so it can be corrected against a real
 * Draco build without touching anything else.
 *
 * Supported feature set (everything the fixtures and scripts/Encoder.py:260
 * "-qp 11 -qt 10 -qn 8 -qg 8 -cl 7" produce, plus the standard traversal):
 *   bitstream 2.2, TRIANGULAR_MESH, MESH_EDGEBREAKER, traversal STANDARD(0)/VALENCE(2),
 *   vertex + corner attribute decoders, depth-first traversal, sequential decoders
 *   INTEGER / QUANTIZATION / NORMALS, prediction NONE / DIFFERENCE / PARALLELOGRAM /
 *   TEX_COORDS_PORTABLE / GEOMETRIC_NORMAL, transforms WRAP and
 *   NORMAL_OCTAHEDRON_CANONICALIZED, symbol schemes TAGGED and RAW.
 * Anything else returns a negative status (mirrors the worker's {type:'error'},
 * DRACOLoader.js:451-455).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "oracle.h"

#define INV (-1)

/* ------------------------------------------------------------------ reader */
typedef struct { const uint8_t *b; size_t n, p; int err; } rd_t;

static uint8_t rd_u8(rd_t *r) { if (r->p + 1 > r->n) { r->err = 1; return 0; } return r->b[r->p++]; }
static int8_t rd_i8(rd_t *r) { return (int8_t)rd_u8(r); }
static uint16_t rd_u16(rd_t *r) { uint16_t a = rd_u8(r); uint16_t b = rd_u8(r); return (uint16_t)(a | (b << 8)); }
static uint32_t rd_u32(rd_t *r) {
    if (r->p + 4 > r->n) { r->err = 1; r->p = r->n; return 0; }
    uint32_t v; memcpy(&v, r->b + r->p, 4); r->p += 4; return v;
}
static int32_t rd_i32(rd_t *r) { return (int32_t)rd_u32(r); }
static float rd_f32(rd_t *r) { uint32_t v = rd_u32(r); float f; memcpy(&f, &v, 4); return f; }
static uint64_t rd_varint(rd_t *r) {
    uint64_t v = 0; int sh = 0;
    for (;;) {
        uint8_t c = rd_u8(r);
        if (r->err) return 0;
        v |= (uint64_t)(c & 0x7f) << sh; sh += 7;
        if (!(c & 0x80)) return v;
        if (sh > 63) { r->err = 1; return 0; }
    }
}

/* ------------------------------------------------------------------ rABS bit stream */
typedef struct { const uint8_t *buf; int off; uint32_t state; uint32_t p; int ok; } rabs_t;

static void rabs_start(rd_t *r, rabs_t *a) {
    memset(a, 0, sizeof *a);
    uint8_t prob_zero = rd_u8(r);
    uint64_t size = rd_varint(r);
    if (r->err || size > r->n - r->p || size == 0) { r->err = 1; return; }
    a->buf = r->b + r->p; a->p = 256u - prob_zero;
    const uint8_t *b = a->buf; int n = (int)size;
    unsigned x = b[n - 1] >> 6;
    if (x == 0) { a->off = n - 1; a->state = b[n - 1] & 0x3f; }
    else if (x == 1) { if (n < 2) { r->err = 1; return; } a->off = n - 2; a->state = (b[n - 2] | (b[n - 1] << 8)) & 0x3fff; }
    else if (x == 2) { if (n < 3) { r->err = 1; return; } a->off = n - 3; a->state = (b[n - 3] | (b[n - 2] << 8) | (b[n - 1] << 16)) & 0x3fffff; }
    else { r->err = 1; return; }
    a->state += 4096u; a->ok = 1;
    r->p += size;
}
static inline int rabs_bit(rabs_t *a) {
    if (a->state < 4096u && a->off > 0) a->state = a->state * 256u + a->buf[--a->off];
    uint32_t x = a->state, q = x >> 8, rem = x & 255u, xn = q * a->p;
    if (rem < a->p) { a->state = xn + rem; return 1; }
    a->state = x - xn - a->p; return 0;
}

/* ------------------------------------------------------------------ rANS symbol decode */
typedef struct { uint32_t n; uint32_t *prob; uint32_t *cum; } ptab_t;

static int ptab_read(rd_t *r, ptab_t *t) {
    uint64_t n = rd_varint(r);
    if (r->err || n > (1u << 20)) return -1;
    t->n = (uint32_t)n;
    t->prob = (uint32_t *)calloc(n + 1, 4); t->cum = (uint32_t *)calloc(n + 2, 4);
    uint32_t i = 0;
    while (i < n) {
        uint8_t d = rd_u8(r); unsigned tok = d & 3;
        if (r->err) return -1;
        if (tok == 3) { uint32_t run = (d >> 2) + 1; if (i + run > n) return -1; i += run; }
        else {
            uint32_t pr = d >> 2;
            for (unsigned k = 0; k < tok; k++) pr |= (uint32_t)rd_u8(r) << (8 * (k + 1) - 2);
            t->prob[i++] = pr;
        }
    }
    for (i = 0; i < n; i++) t->cum[i + 1] = t->cum[i] + t->prob[i];
    return r->err ? -1 : 0;
}
static void ptab_free(ptab_t *t) { free(t->prob); free(t->cum); t->prob = t->cum = NULL; }

/* decode `count` symbols with table t at precision pb; the byte run is length-prefixed (varint).
 * term_ok (optional) reports the A.4 terminal-state oracle: state==l_base and no bytes left. */
static int rans_decode(rd_t *r, const ptab_t *t, int pb, uint64_t count, uint32_t *out, int *term_ok) {
    uint64_t nbytes = rd_varint(r);
    if (r->err || nbytes > r->n - r->p) return -1;
    const uint8_t *b = r->b + r->p; r->p += nbytes;
    if (term_ok) *term_ok = 1;
    if (count == 0) return 0;
    if (nbytes == 0 || t->n == 0) return -1;
    uint32_t prec = 1u << pb, lbase = prec * 4u;
    if (t->cum[t->n] != prec) return -1;
    unsigned x = b[nbytes - 1] >> 6; unsigned k = x + 1;
    if (nbytes < k) return -1;
    uint32_t st = 0;
    for (unsigned i = 0; i < k; i++) st |= (uint32_t)b[nbytes - k + i] << (8 * i);
    st &= (1u << (8 * k - 2)) - 1u;
    int64_t off = (int64_t)nbytes - k;
    st += lbase;
    uint32_t *lut = (uint32_t *)malloc((size_t)prec * 4);
    for (uint32_t s = 0; s < t->n; s++) for (uint32_t j = t->cum[s]; j < t->cum[s + 1]; j++) lut[j] = s;
    for (uint64_t i = 0; i < count; i++) {
        while (st < lbase && off > 0) st = st * 256u + b[--off];
        uint32_t q = st >> pb, rem = st & (prec - 1);
        uint32_t s = lut[rem];
        st = q * t->prob[s] + rem - t->cum[s];
        out[i] = s;
    }
    free(lut);
    /* A.4 terminal-state oracle: after one last renormalisation the state must be back at
     * l_base with the byte run exhausted (the encoder started from l_base). */
    while (st < lbase && off > 0) st = st * 256u + b[--off];
    if (term_ok) *term_ok = (st == lbase && off == 0);
    return 0;
}

/* DecodeSymbols: scheme u8; RAW(1) or TAGGED(0). */
static int decode_symbols(rd_t *r, uint64_t num_values, int nc, uint32_t *out, int *term_ok) {
    uint8_t scheme = rd_u8(r);
    if (r->err) return -1;
    if (scheme == 1) {
        uint8_t mbl = rd_u8(r);
        ptab_t t = {0};
        if (ptab_read(r, &t)) { ptab_free(&t); return -1; }
        int pb = (3 * mbl) / 2; if (pb < 12) pb = 12; if (pb > 20) pb = 20;
        int rc = rans_decode(r, &t, pb, num_values, out, term_ok);
        ptab_free(&t);
        return rc;
    } else if (scheme == 0) {
        ptab_t t = {0};
        if (ptab_read(r, &t)) { ptab_free(&t); return -1; }
        uint64_t ntags = num_values / (uint64_t)nc;
        uint32_t *tags = (uint32_t *)malloc((ntags + 1) * 4);
        int rc = rans_decode(r, &t, 12, ntags, tags, term_ok);
        ptab_free(&t);
        if (rc) { free(tags); return rc; }
        /* raw LSB-first bit fields, byte padded */
        uint64_t bp = 0; const uint8_t *b = r->b + r->p; size_t avail = r->n - r->p;
        for (uint64_t i = 0; i < ntags; i++) {
            uint32_t bl = tags[i];
            for (int c = 0; c < nc; c++) {
                uint32_t v = 0;
                for (uint32_t k = 0; k < bl; k++) {
                    uint64_t byte = bp >> 3;
                    if (byte >= avail) { free(tags); return -1; }
                    v |= (uint32_t)((b[byte] >> (bp & 7)) & 1) << k; bp++;
                }
                out[i * nc + c] = v;
            }
        }
        r->p += (bp + 7) >> 3;
        free(tags);
        return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------ corner helpers */
static inline int c_next(int c) { return c < 0 ? INV : ((c % 3) == 2 ? c - 2 : c + 1); }
static inline int c_prev(int c) { return c < 0 ? INV : ((c % 3) == 0 ? c + 2 : c - 1); }

typedef struct {
    int F, V;            /* faces, vertex slots (incl. isolated tail) */
    int *opp, *c2v;      /* [3F] */
    int *lmc;            /* [V] left-most corner, INV if isolated */
    uint8_t *hole;       /* [V] */
} ctab_t;

static inline int b_swl(const ctab_t *t, int c) { if (c < 0) return INV; int o = t->opp[c_next(c)]; return o < 0 ? INV : c_next(o); }
static inline int b_swr(const ctab_t *t, int c) { if (c < 0) return INV; int o = t->opp[c_prev(c)]; return o < 0 ? INV : c_prev(o); }

/* attribute corner table view: same faces, seams cut `Opposite`. */
typedef struct {
    const ctab_t *base;
    uint8_t *eos;        /* [3F] corner is opposite a seam edge */
    uint8_t *vos;        /* [V]  base vertex touches a seam edge */
    int *c2v;            /* [3F] attribute vertex id */
    int *lmc; int nv;    /* [nv] left-most corner per attribute vertex */
} atab_t;

static inline int a_opp(const atab_t *a, int c) { if (c < 0 || a->eos[c]) return INV; return a->base->opp[c]; }
static inline int a_swl(const atab_t *a, int c) { if (c < 0) return INV; int o = a_opp(a, c_next(c)); return o < 0 ? INV : c_next(o); }
static inline int a_swr(const atab_t *a, int c) { if (c < 0) return INV; int o = a_opp(a, c_prev(c)); return o < 0 ? INV : c_prev(o); }

/* generic view used by traversal/prediction: att==NULL -> base table */
typedef struct { const ctab_t *b; const atab_t *a; } view_t;
static inline int v_opp(const view_t *v, int c) { return v->a ? a_opp(v->a, c) : (c < 0 ? INV : v->b->opp[c]); }
static inline int v_vert(const view_t *v, int c) { return v->a ? v->a->c2v[c] : v->b->c2v[c]; }
static inline int v_swl(const view_t *v, int c) { return v->a ? a_swl(v->a, c) : b_swl(v->b, c); }
static inline int v_swr(const view_t *v, int c) { return v->a ? a_swr(v->a, c) : b_swr(v->b, c); }
static inline int v_nverts(const view_t *v) { return v->a ? v->a->nv : v->b->V; }
static inline int v_on_boundary(const view_t *v, int vert) {
    if (v->a) { int c = v->a->lmc[vert]; if (c < 0) return 1; return v->a->vos[v->b->c2v[c]]; }
    int c = v->b->lmc[vert]; return b_swl(v->b, c) == INV;
}

/* ------------------------------------------------------------------ connectivity */
typedef struct { uint32_t src, split; uint8_t edge; } tsplit_t;

typedef struct {
    rd_t r;
    int trav, nv_enc, nf, nad, nsym, nsplit;
    tsplit_t *ts; int nts;
    rabs_t start_faces; rabs_t seams[8];
    uint32_t *ctx[6]; int ctx_n[6];
    /* standard traversal symbol bit buffer */
    const uint8_t *sym_buf; size_t sym_len; uint64_t sym_bp;
    ctab_t t;
    int symhist[5];
    int ctx_left_nonzero;
} conn_t;

static int std_symbol(conn_t *c) {
    /* LSB-first: 1 bit; 0 -> C ; else 2 more bits, symbol = 1 | (suffix << 1): S=1 L=3 R=5 E=7 */
    uint64_t bp = c->sym_bp;
    if ((bp >> 3) >= c->sym_len) return -1;
    int b0 = (c->sym_buf[bp >> 3] >> (bp & 7)) & 1; bp++;
    if (!b0) { c->sym_bp = bp; return 0; }
    int suf = 0;
    for (int k = 0; k < 2; k++) { if ((bp >> 3) >= c->sym_len) return -1; suf |= ((c->sym_buf[bp >> 3] >> (bp & 7)) & 1) << k; bp++; }
    c->sym_bp = bp;
    /* topology ids: C0 S1 L2 R3 E4 */
    static const int map[4] = {1, 2, 3, 4};
    return map[suf];
}

static void set_opp(ctab_t *t, int a, int b) { t->opp[a] = b; t->opp[b] = a; }

static int decode_connectivity(conn_t *cn) {
    rd_t *r = &cn->r;
    cn->trav = rd_u8(r);
    cn->nv_enc = (int)rd_varint(r); cn->nf = (int)rd_varint(r); cn->nad = rd_u8(r);
    cn->nsym = (int)rd_varint(r); cn->nsplit = (int)rd_varint(r);
    if (r->err) return UVO_ERR_TRUNCATED;
    if (cn->trav != 0 && cn->trav != 2) return UVO_ERR_UNSUPPORTED;
    if (cn->nad > 8 || cn->nf <= 0 || cn->nf > (1 << 26) || cn->nsym > cn->nf || cn->nv_enc > 3 * cn->nf + 3) return UVO_ERR_CORRUPT;
    /* topology split events */
    cn->nts = (int)rd_varint(r);
    if (r->err || cn->nts > cn->nf) return UVO_ERR_CORRUPT;
    cn->ts = (tsplit_t *)calloc((size_t)cn->nts + 1, sizeof(tsplit_t));
    uint32_t last = 0;
    for (int i = 0; i < cn->nts; i++) {
        uint32_t d = (uint32_t)rd_varint(r); uint32_t src = last + d; uint32_t d2 = (uint32_t)rd_varint(r);
        cn->ts[i].src = src; cn->ts[i].split = src - d2; last = src;
    }
    if (cn->nts > 0) {
        size_t nb = ((size_t)cn->nts + 7) / 8;
        if (r->p + nb > r->n) return UVO_ERR_TRUNCATED;
        for (int i = 0; i < cn->nts; i++) cn->ts[i].edge = (r->b[r->p + (i >> 3)] >> (i & 7)) & 1;
        r->p += nb;
    }
    if (cn->trav == 0) {
        uint64_t sz = rd_varint(r);
        if (r->err || sz > r->n - r->p) return UVO_ERR_TRUNCATED;
        cn->sym_buf = r->b + r->p; cn->sym_len = sz; cn->sym_bp = 0; r->p += sz;
    }
    rabs_start(r, &cn->start_faces);
    for (int i = 0; i < cn->nad; i++) rabs_start(r, &cn->seams[i]);
    if (r->err) return UVO_ERR_TRUNCATED;
    if (cn->trav == 2) {
        for (int i = 0; i < 6; i++) {
            uint64_t n = rd_varint(r);
            if (r->err || n > (uint64_t)cn->nsym) return UVO_ERR_CORRUPT;
            cn->ctx_n[i] = (int)n;
            if (n) {
                cn->ctx[i] = (uint32_t *)malloc(n * 4);
                if (decode_symbols(r, n, 1, cn->ctx[i], NULL)) return UVO_ERR_CORRUPT;
            }
        }
    }
    const int F = cn->nf, maxv = cn->nv_enc + cn->nsplit;
    ctab_t *t = &cn->t;
    t->F = F;
    t->opp = (int *)malloc((size_t)3 * F * 4); t->c2v = (int *)malloc((size_t)3 * F * 4);
    for (int i = 0; i < 3 * F; i++) { t->opp[i] = INV; t->c2v[i] = INV; }
    t->lmc = (int *)malloc(((size_t)maxv + 4) * 4); t->hole = (uint8_t *)malloc((size_t)maxv + 4);
    int *val = (int *)calloc((size_t)maxv + 4, 4);
    memset(t->hole, 1, (size_t)maxv + 4);
    int nverts = 0, numf = 0;
    int *stack = (int *)malloc(((size_t)cn->nsym + 8) * 4); int sp = 0;
    int *invalid = (int *)malloc(((size_t)cn->nsplit + 8) * 4); int ninv = 0;
    /* split_active: symbol id -> corner pushed when that S is reached */
    int *split_key = (int *)malloc(((size_t)cn->nts + 1) * 4), *split_val = (int *)malloc(((size_t)cn->nts + 1) * 4); int nsa = 0;
    int ts_top = cn->nts;  /* events consumed from the back */
    int active_ctx = -1, rc = 0;
#define FAIL(code) do { rc = (code); goto done; } while (0)
#define ADDV(dst) do { if (nverts >= maxv) FAIL(UVO_ERR_CORRUPT); t->lmc[nverts] = INV; (dst) = nverts++; } while (0)
    for (int sid = 0; sid < cn->nsym; sid++) {
        if (numf >= F) FAIL(UVO_ERR_CORRUPT);
        int face = numf++; int c0 = 3 * face; int s, chk = 0;
        if (cn->trav == 2) {
            if (active_ctx >= 0) {
                if (cn->ctx_n[active_ctx] <= 0) FAIL(UVO_ERR_CORRUPT);
                s = (int)cn->ctx[active_ctx][--cn->ctx_n[active_ctx]];
                if (s > 4) FAIL(UVO_ERR_CORRUPT);
            } else s = 4;
        } else { s = std_symbol(cn); if (s < 0) FAIL(UVO_ERR_TRUNCATED); }
        cn->symhist[s]++;
        if (s == 0) { /* C */
            if (sp == 0) FAIL(UVO_ERR_CORRUPT);
            int a = stack[sp - 1]; int vx = t->c2v[c_next(a)];
            if (vx < 0 || t->lmc[vx] < 0) FAIL(UVO_ERR_CORRUPT);
            int b = c_next(t->lmc[vx]);
            if (a == b || t->opp[a] >= 0 || t->opp[b] >= 0) FAIL(UVO_ERR_CORRUPT);
            set_opp(t, a, c0 + 1); set_opp(t, b, c0 + 2);
            int vap = t->c2v[c_prev(a)], vbn = t->c2v[c_next(b)];
            if (vx == vap || vx == vbn) FAIL(UVO_ERR_CORRUPT);
            t->c2v[c0] = vx; t->c2v[c0 + 1] = vbn; t->c2v[c0 + 2] = vap;
            t->lmc[vap] = c0 + 2; t->hole[vx] = 0; stack[sp - 1] = c0;
        } else if (s == 3 || s == 2) { /* R / L */
            if (sp == 0) FAIL(UVO_ERR_CORRUPT);
            int a = stack[sp - 1]; if (t->opp[a] >= 0) FAIL(UVO_ERR_CORRUPT);
            int oc, cl, cr;
            if (s == 3) { oc = c0 + 2; cl = c0 + 1; cr = c0; } else { oc = c0 + 1; cl = c0; cr = c0 + 2; }
            set_opp(t, oc, a);
            int nvx; ADDV(nvx);
            t->c2v[oc] = nvx; t->lmc[nvx] = oc;
            int vr = t->c2v[c_prev(a)]; t->c2v[cr] = vr; t->lmc[vr] = cr;
            t->c2v[cl] = t->c2v[c_next(a)];
            stack[sp - 1] = c0; chk = 1;
        } else if (s == 1) { /* S */
            if (sp == 0) FAIL(UVO_ERR_CORRUPT);
            int b = stack[--sp];
            for (int k = 0; k < nsa; k++) if (split_key[k] == sid) { stack[sp++] = split_val[k]; break; }
            if (sp == 0) FAIL(UVO_ERR_CORRUPT);
            int a = stack[sp - 1];
            if (a == b || t->opp[a] >= 0 || t->opp[b] >= 0) FAIL(UVO_ERR_CORRUPT);
            set_opp(t, a, c0 + 2); set_opp(t, b, c0 + 1);
            int vp = t->c2v[c_prev(a)];
            t->c2v[c0] = vp; t->c2v[c0 + 1] = t->c2v[c_next(a)];
            int vbp = t->c2v[c_prev(b)]; t->c2v[c0 + 2] = vbp; t->lmc[vbp] = c0 + 2;
            int cnr = c_next(b); int vn = t->c2v[cnr];
            val[vp] += val[vn];
            t->lmc[vp] = t->lmc[vn];
            int first = cnr;
            while (cnr >= 0) { t->c2v[cnr] = vp; cnr = b_swl(t, cnr); if (cnr == first) FAIL(UVO_ERR_CORRUPT); }
            t->lmc[vn] = INV; invalid[ninv++] = vn;
            stack[sp - 1] = c0;
        } else { /* E */
            int v0, v1, v2; ADDV(v0); ADDV(v1); ADDV(v2);
            t->c2v[c0] = v0; t->c2v[c0 + 1] = v1; t->c2v[c0 + 2] = v2;
            t->lmc[v0] = c0; t->lmc[v1] = c0 + 1; t->lmc[v2] = c0 + 2;
            stack[sp++] = c0; chk = 1;
        }
        if (cn->trav == 2) { /* valence contexts */
            int c = stack[sp - 1], n_ = c_next(c), p_ = c_prev(c);
            if (s == 0 || s == 1) { val[t->c2v[n_]] += 1; val[t->c2v[p_]] += 1; }
            else if (s == 3) { val[t->c2v[c]] += 1; val[t->c2v[n_]] += 1; val[t->c2v[p_]] += 2; }
            else if (s == 2) { val[t->c2v[c]] += 1; val[t->c2v[n_]] += 2; val[t->c2v[p_]] += 1; }
            else { val[t->c2v[c]] += 2; val[t->c2v[n_]] += 2; val[t->c2v[p_]] += 2; }
            int v = val[t->c2v[n_]]; if (v < 2) v = 2; if (v > 7) v = 7;
            active_ctx = v - 2;
        }
        if (chk) {
            uint32_t enc_id = (uint32_t)(cn->nsym - sid - 1);
            while (ts_top > 0 && cn->ts[ts_top - 1].src == enc_id) {
                tsplit_t e = cn->ts[--ts_top];
                int top = stack[sp - 1];
                int nac = e.edge == 1 ? c_next(top) : c_prev(top);
                split_key[nsa] = cn->nsym - (int)e.split - 1; split_val[nsa] = nac; nsa++;
            }
        }
    }
    /* start faces */
    while (sp > 0) {
        int corner = stack[--sp];
        if (rabs_bit(&cn->start_faces)) {
            int a = corner; int vn = t->c2v[c_next(a)];
            if (vn < 0 || t->lmc[vn] < 0) FAIL(UVO_ERR_CORRUPT);
            int cb = c_next(t->lmc[vn]); int vx = t->c2v[c_next(cb)];
            if (vx < 0 || t->lmc[vx] < 0) FAIL(UVO_ERR_CORRUPT);
            int cc = c_next(t->lmc[vx]);
            if (a == cb || cb == cc || a == cc || t->opp[a] >= 0 || t->opp[cb] >= 0 || t->opp[cc] >= 0) FAIL(UVO_ERR_CORRUPT);
            int vp = t->c2v[c_next(cc)];
            if (numf >= F) FAIL(UVO_ERR_CORRUPT);
            int nc = 3 * numf++;
            set_opp(t, nc, a); set_opp(t, nc + 1, cb); set_opp(t, nc + 2, cc);
            t->c2v[nc] = vx; t->c2v[nc + 1] = vp; t->c2v[nc + 2] = vn;
            for (int k = 0; k < 3; k++) t->hole[t->c2v[nc + k]] = 0;
        }
    }
    if (numf != F) FAIL(UVO_ERR_CORRUPT);
    cn->ctx_left_nonzero = 0;
    for (int i = 0; i < 6; i++) if (cn->ctx_n[i] != 0) cn->ctx_left_nonzero = 1;
    /* Remove isolated vertices by moving the last valid vertex into each hole
     * (upstream DecodeConnectivity tail; defines final vertex ids and hence point order). */
    {
        int num_vertices = nverts;
        for (int k = 0; k < ninv; k++) {
            int iv = invalid[k];
            int src = num_vertices - 1;
            while (src >= 0 && t->lmc[src] == INV) src = --num_vertices - 1;
            if (src < iv) continue;
            /* remap all corners of src to iv: iterate the full fan (left then right) */
            int c0 = t->lmc[src], c = c0, left = 1;
            while (c >= 0) {
                int nx;
                if (left) { nx = b_swl(t, c); if (nx < 0) { nx = b_swr(t, c0); left = 0; } else if (nx == c0) nx = INV; }
                else nx = b_swr(t, c);
                if (t->c2v[c] != src) FAIL(UVO_ERR_CORRUPT);
                t->c2v[c] = iv; c = nx;
            }
            t->lmc[iv] = t->lmc[src]; t->lmc[src] = INV;
            t->hole[iv] = t->hole[src]; t->hole[src] = 0;
            num_vertices--;
        }
        t->V = nverts;  /* slots; isolated ones are now all at the tail with lmc==INV */
    }
done:
    free(val); free(stack); free(invalid); free(split_key); free(split_val);
    return rc;
#undef FAIL
#undef ADDV
}

/* attribute seams + attribute corner tables */
static int decode_attribute_connectivity(conn_t *cn, atab_t *at) {
    const ctab_t *t = &cn->t; const int F = t->F, V = t->V, nad = cn->nad;
    for (int i = 0; i < nad; i++) {
        at[i].base = t;
        at[i].eos = (uint8_t *)calloc((size_t)3 * F, 1); at[i].vos = (uint8_t *)calloc((size_t)V + 1, 1);
        at[i].c2v = (int *)malloc((size_t)3 * F * 4);
        for (int c = 0; c < 3 * F; c++) at[i].c2v[c] = INV;
        at[i].lmc = (int *)malloc((size_t)3 * F * 4); at[i].nv = 0;
    }
#define ADD_SEAM(i, c) do { int c_ = (c); at[i].eos[c_] = 1; at[i].vos[t->c2v[c_next(c_)]] = 1; at[i].vos[t->c2v[c_prev(c_)]] = 1; \
        int o_ = t->opp[c_]; if (o_ >= 0) { at[i].eos[o_] = 1; at[i].vos[t->c2v[c_next(o_)]] = 1; at[i].vos[t->c2v[c_prev(o_)]] = 1; } } while (0)
    for (int f = 0; f < F; f++) {
        for (int k = 0; k < 3; k++) {
            int c = 3 * f + k, o = t->opp[c];
            if (o < 0) { for (int i = 0; i < nad; i++) ADD_SEAM(i, c); continue; }
            if (o / 3 < f) continue;
            for (int i = 0; i < nad; i++) if (rabs_bit(&cn->seams[i])) ADD_SEAM(i, c);
        }
    }
#undef ADD_SEAM
    for (int i = 0; i < nad; i++) {
        atab_t *a = &at[i]; int nnew = 0;
        for (int v = 0; v < V; v++) {
            int c = t->lmc[v];
            if (c < 0) continue;
            int fid = nnew++, fc = c;
            if (a->vos[v]) {
                int x = a_swl(a, fc);
                while (x >= 0) { fc = x; x = a_swl(a, x); if (x == c) return UVO_ERR_CORRUPT; }
            }
            a->c2v[fc] = fid; a->lmc[fid] = fc;
            int x = b_swr(t, fc);
            while (x >= 0 && x != fc) {
                if (a->eos[c_next(x)]) { fid = nnew++; a->lmc[fid] = x; }
                a->c2v[x] = fid; x = b_swr(t, x);
            }
        }
        a->nv = nnew;
    }
    return 0;
}

/* AssignPointsToCorners (unpinned ordering isolated here) */
static int assign_points(const conn_t *cn, const atab_t *at, int *c2p, int *p2c) {
    const ctab_t *t = &cn->t; const int V = t->V, nad = cn->nad; int np = 0;
    if (nad == 0) {
        /* identity: point id == vertex id */
        for (int c = 0; c < 3 * t->F; c++) c2p[c] = t->c2v[c];
        int n = 0; for (int v = 0; v < V; v++) if (t->lmc[v] >= 0) { p2c[v] = t->lmc[v]; n = v + 1; }
        return n;
    }
    for (int v = 0; v < V; v++) {
        int c = t->lmc[v];
        if (c < 0) continue;
        int dfc = c;
        if (!t->hole[v]) {
            for (int i = 0; i < nad; i++) {
                if (!at[i].vos[t->c2v[c]]) continue;
                int vid = at[i].c2v[c], x = b_swr(t, c), found = 0;
                while (x != c) {
                    if (x < 0) return -1;
                    if (at[i].c2v[x] != vid) { dfc = x; found = 1; break; }
                    x = b_swr(t, x);
                }
                if (found) break;
            }
        }
        c = dfc; c2p[c] = np; p2c[np++] = c;
        int pc = c; c = b_swr(t, c);
        while (c >= 0 && c != dfc) {
            int seam = 0;
            for (int i = 0; i < nad; i++) if (at[i].c2v[c] != at[i].c2v[pc]) { seam = 1; break; }
            if (seam) { c2p[c] = np; p2c[np++] = c; } else c2p[c] = c2p[pc];
            pc = c; c = b_swr(t, c);
        }
    }
    return np;
}

/* depth-first traversal -> entry order (SURVEY A.3 "Traversal / sequencing") */
static int traverse(const view_t *vw, int *d2c, int *v2d) {
    const int F = vw->b->F, nv = v_nverts(vw);
    uint8_t *fvis = (uint8_t *)calloc((size_t)F + 1, 1), *vvis = (uint8_t *)calloc((size_t)nv + 1, 1);
    int *st = (int *)malloc(((size_t)F + 4) * 4); int n = 0, rc = 0;
    for (int i = 0; i < nv; i++) v2d[i] = -1;
#define VISIT(v, c) do { vvis[v] = 1; v2d[v] = n; d2c[n++] = (c); } while (0)
#define FVIS(c) ((c) < 0 ? 1 : fvis[(c) / 3])
    for (int f = 0; f < F; f++) {
        int c = 3 * f;
        if (fvis[f]) continue;
        int sp = 0; st[sp++] = c;
        int nvx = v_vert(vw, c_next(c)), pvx = v_vert(vw, c_prev(c));
        if (nvx < 0 || pvx < 0) { rc = -1; goto out; }
        if (!vvis[nvx]) VISIT(nvx, c_next(c));
        if (!vvis[pvx]) VISIT(pvx, c_prev(c));
        while (sp > 0) {
            c = st[sp - 1];
            if (c < 0 || fvis[c / 3]) { sp--; continue; }
            for (;;) {
                fvis[c / 3] = 1;
                int v = v_vert(vw, c);
                if (v < 0) { rc = -1; goto out; }
                if (!vvis[v]) {
                    int ob = v_on_boundary(vw, v);
                    VISIT(v, c);
                    if (!ob) { c = v_opp(vw, c_next(c)); if (c < 0) { rc = -1; goto out; } continue; }
                }
                int rcn = v_opp(vw, c_next(c)), lcn = v_opp(vw, c_prev(c));
                if (FVIS(rcn)) {
                    if (FVIS(lcn)) { sp--; break; }
                    c = lcn;
                } else {
                    if (FVIS(lcn)) c = rcn;
                    else { st[sp - 1] = lcn; st[sp++] = rcn; break; }
                }
            }
        }
    }
#undef VISIT
#undef FVIS
out:
    free(fvis); free(vvis); free(st);
    return rc ? rc : n;
}

/* ------------------------------------------------------------------ prediction */
static inline int32_t zigzag(uint32_t s) { return (s & 1) ? -(int32_t)(s >> 1) - 1 : (int32_t)(s >> 1); }

static inline void wrap_add(const int64_t *pred, const int32_t *corr, int nc, int32_t mn, int32_t mx, int32_t *out, int *nwrap) {
    int32_t md = 1 + mx - mn;
    for (int k = 0; k < nc; k++) {
        int64_t p = pred[k]; if (p > mx) p = mx; if (p < mn) p = mn;
        int32_t o = (int32_t)p + corr[k];
        if (o > mx) { o -= md; if (nwrap) (*nwrap)++; } else if (o < mn) { o += md; if (nwrap) (*nwrap)++; }
        out[k] = o;
    }
}

static uint64_t int_sqrt(uint64_t n) {
    if (n == 0) return 0;
    uint64_t a = n, r = 1;
    while (a >= 2) { r *= 2; a /= 4; }
    do { r = (r + n / r) / 2; } while (r * r > n);
    return r;
}

typedef struct {
    int type, dtype, nc, normalized, unique_id, seq;   /* as in stream */
    int pred, xform;
    int n;                /* entries */
    int32_t *val;         /* [n * vnc] portable ints */
    int vnc;              /* portable components (2 for NORMALS) */
    int *d2c, *v2d;       /* entry->corner, (attr)vertex->entry ; shared per decoder */
    int32_t wmin, wmax;
    float qmin[4], qrange; int qbits;
    int nwrap, term_ok, orient_left, nflip, npar;
} attr_t;

typedef struct { int att_data_id, dec_type, trav_method, natt; attr_t att[8]; int n; int *d2c, *v2d; } adec_t;

/* ------------------------------------------------------------------ fp32 tails (unpinned; isolated) */
static inline float dequant_value(int32_t q, float delta, float mn) {
    /* two separately rounded fp32 ops; the Makefile builds with -ffp-contract=off */
    float a = (float)q * delta;
    return a + mn;
}
static void oct_to_unit(int32_t s, int32_t t, int32_t max_v, float *out) {
    const float scale = 2.0f / (float)max_v;
    float y = (float)s * scale - 1.0f, z = (float)t * scale - 1.0f;
    const float x = 1.0f - fabsf(y) - fabsf(z);
    float xo = -x; xo = xo < 0 ? 0 : xo;
    y += y < 0 ? xo : -xo; z += z < 0 ? xo : -xo;
    const float n2 = x * x + y * y + z * z;
    if (n2 < 1e-6f) { out[0] = out[1] = out[2] = 0; return; }
    const float d = 1.0f / sqrtf(n2);
    out[0] = x * d; out[1] = y * d; out[2] = z * d;
}

/* ------------------------------------------------------------------ main entry */
static void free_atab(atab_t *a) { free(a->eos); free(a->vos); free(a->c2v); free(a->lmc); }

void uvo_draco_free(uvo_draco_mesh *m) {
    if (!m) return;
    free(m->index); free(m->position); free(m->normal); free(m->uv); free(m->color);
    free(m->pos_q); free(m->uv_q); free(m->nrm_q); free(m->dbg_c2v); free(m->dbg_opp);
    memset(m, 0, sizeof *m);
}

/* Metadata section (header flag 0x8000; written by draco::MetadataEncoder::EncodeGeometryMetadata): decoded past, never used --
 * DRACOLoader reads no metadata (src/lib/DRACOLoader.js:470-554).  Layout: varint number of attribute metadata, each {varint attribute
 * unique id, metadata}; then the geometry's own metadata.  metadata = varint entries, each {u8 name length, name, varint value size,
 * value}; varint sub-metadata count, each {u8 name length, name, metadata}. */
static int skip_metadata(rd_t *r, int depth) {
    if (depth > 32) return -1;
    uint64_t ne = rd_varint(r);
    if (r->err || ne > r->n) return -1;
    for (uint64_t i = 0; i < ne; i++) {
        uint8_t nl = rd_u8(r); if (r->err || nl > r->n - r->p) return -1; r->p += nl;
        uint64_t vs = rd_varint(r); if (r->err || vs > r->n - r->p) return -1; r->p += vs;
    }
    uint64_t ns = rd_varint(r);
    if (r->err || ns > r->n) return -1;
    for (uint64_t i = 0; i < ns; i++) {
        uint8_t nl = rd_u8(r); if (r->err || nl > r->n - r->p) return -1; r->p += nl;
        if (skip_metadata(r, depth + 1)) return -1;
    }
    return 0;
}

int uvo_draco_decode(const uint8_t *data, size_t len, uvo_draco_mesh *out) {
    memset(out, 0, sizeof *out);
    conn_t cn; memset(&cn, 0, sizeof cn);
    atab_t at[8]; memset(at, 0, sizeof at);
    adec_t dec[8]; memset(dec, 0, sizeof dec);
    int ndec = 0, rc = 0;
    int *c2p = NULL, *p2c = NULL;
    rd_t *r = &cn.r; r->b = data; r->n = len; r->p = 0;
#define FAIL(code) do { rc = (code); goto done; } while (0)
    if (len < 11 || memcmp(data, "DRACO", 5)) FAIL(UVO_ERR_CORRUPT);
    r->p = 5;
    int maj = rd_u8(r), mino = rd_u8(r), etype = rd_u8(r), meth = rd_u8(r); int flags = rd_u16(r);
    if (maj != 2 || mino != 2) FAIL(UVO_ERR_UNSUPPORTED);
    if (etype != 1 || meth != 1) FAIL(UVO_ERR_UNSUPPORTED);
    if (flags & 0x8000) {
        uint64_t na = rd_varint(r);
        if (r->err || na > len) FAIL(UVO_ERR_CORRUPT);
        for (uint64_t i = 0; i < na; i++) { (void)rd_varint(r); if (skip_metadata(r, 0)) FAIL(UVO_ERR_CORRUPT); }
        if (skip_metadata(r, 0)) FAIL(UVO_ERR_CORRUPT);
    }
    if ((rc = decode_connectivity(&cn))) goto done;
    if ((rc = decode_attribute_connectivity(&cn, at))) goto done;
    const ctab_t *t = &cn.t; const int F = t->F, nad = cn.nad;
    c2p = (int *)malloc((size_t)3 * F * 4); p2c = (int *)malloc(((size_t)3 * F + t->V + 4) * 4);
    int np = assign_points(&cn, at, c2p, p2c);
    if (np < 0) FAIL(UVO_ERR_CORRUPT);
    out->num_faces = (uint32_t)F; out->num_points = (uint32_t)np;
    out->num_vertices = (uint32_t)cn.nv_enc; out->num_symbols = (uint32_t)cn.nsym;
    for (int i = 0; i < 5; i++) out->symhist[i] = (uint32_t)cn.symhist[i];
    out->ctx_counters_zero = !cn.ctx_left_nonzero;
    for (int i = 0; i < nad && i < 4; i++) out->attr_vertices[i] = (uint32_t)at[i].nv;
    out->index = (uint32_t *)malloc((size_t)3 * F * 4);
    for (int c = 0; c < 3 * F; c++) out->index[c] = (uint32_t)c2p[c];

    /* ---- attribute decoders: headers */
    ndec = rd_u8(r);
    if (r->err || ndec > 8) FAIL(UVO_ERR_CORRUPT);
    for (int i = 0; i < ndec; i++) { dec[i].att_data_id = rd_i8(r); dec[i].dec_type = rd_u8(r); dec[i].trav_method = rd_u8(r); }
    for (int i = 0; i < ndec; i++) {
        adec_t *d = &dec[i];
        if (d->att_data_id >= nad || d->trav_method != 0 || d->dec_type > 1) FAIL(UVO_ERR_UNSUPPORTED);
        if (d->dec_type == 1 && d->att_data_id < 0) FAIL(UVO_ERR_CORRUPT);
        d->natt = (int)rd_varint(r);
        if (r->err || d->natt > 8 || d->natt < 1) FAIL(UVO_ERR_CORRUPT);
        for (int j = 0; j < d->natt; j++) {
            attr_t *a = &d->att[j];
            a->type = rd_u8(r); a->dtype = rd_u8(r); a->nc = rd_u8(r); a->normalized = rd_u8(r); a->unique_id = (int)rd_varint(r);
            if (a->nc < 1 || a->nc > 4) FAIL(UVO_ERR_UNSUPPORTED);
        }
        for (int j = 0; j < d->natt; j++) d->att[j].seq = rd_u8(r);
    }
    if (r->err) FAIL(UVO_ERR_TRUNCATED);

    /* ---- per decoder: traversal, portable data, transform data */
    const adec_t *posdec = NULL; const attr_t *posatt = NULL;
    for (int i = 0; i < ndec; i++) {
        adec_t *d = &dec[i];
        view_t vw; vw.b = t; vw.a = (d->dec_type == 1) ? &at[d->att_data_id] : NULL;
        int nvv = v_nverts(&vw);
        d->d2c = (int *)malloc(((size_t)nvv + 4) * 4); d->v2d = (int *)malloc(((size_t)nvv + 4) * 4);
        int n = traverse(&vw, d->d2c, d->v2d);
        if (n < 0) FAIL(UVO_ERR_CORRUPT);
        d->n = n;
        for (int j = 0; j < d->natt; j++) {
            attr_t *a = &d->att[j];
            a->n = n; a->d2c = d->d2c; a->v2d = d->v2d;
            a->vnc = (a->seq == 3) ? 2 : a->nc;
            if (a->seq < 1 || a->seq > 3) FAIL(UVO_ERR_UNSUPPORTED);
            a->pred = rd_i8(r); a->xform = -1;
            if (a->pred != -2) a->xform = rd_i8(r);
            int compressed = rd_u8(r);
            if (r->err) FAIL(UVO_ERR_TRUNCATED);
            size_t nvals = (size_t)n * a->vnc;
            uint32_t *sym = (uint32_t *)malloc((nvals + 4) * 4);
            a->val = (int32_t *)malloc((nvals + 4) * 4);
            a->term_ok = 1;
            if (compressed) {
                if (decode_symbols(r, nvals, a->vnc, sym, &a->term_ok)) { free(sym); FAIL(UVO_ERR_CORRUPT); }
            } else {
                int nb = rd_u8(r);
                if (nb < 1 || nb > 4 || r->p + nvals * nb > r->n) { free(sym); FAIL(UVO_ERR_CORRUPT); }
                for (size_t k = 0; k < nvals; k++) { uint32_t v = 0; memcpy(&v, r->b + r->p, nb); r->p += nb; sym[k] = v; }
            }
            int positive = (a->pred != -2) && (a->xform == 2 || a->xform == 3);
            int32_t *corr = (int32_t *)sym;
            if (!positive) for (size_t k = 0; k < nvals; k++) corr[k] = zigzag(sym[k]);
            const int nc = a->vnc;
            /* prediction data + reversal */
            if (a->pred == -2) {
                memcpy(a->val, corr, nvals * 4);
            } else if (a->pred == 0 || a->pred == 1) { /* DIFFERENCE / PARALLELOGRAM, WRAP */
                if (a->xform != 1) { free(sym); FAIL(UVO_ERR_UNSUPPORTED); }
                a->wmin = rd_i32(r); a->wmax = rd_i32(r);
                if (r->err || a->wmax < a->wmin) { free(sym); FAIL(UVO_ERR_CORRUPT); }
                int64_t pred[4] = {0, 0, 0, 0};
                if (n > 0) wrap_add(pred, corr, nc, a->wmin, a->wmax, a->val, &a->nwrap);
                for (int p = 1; p < n; p++) {
                    int ok = 0;
                    if (a->pred == 1) {
                        int ci = d->d2c[p], oci = v_opp(&vw, ci);
                        if (oci >= 0) {
                            int eo = d->v2d[v_vert(&vw, oci)], en = d->v2d[v_vert(&vw, c_next(oci))], ep = d->v2d[v_vert(&vw, c_prev(oci))];
                            if (eo < p && en < p && ep < p && eo >= 0 && en >= 0 && ep >= 0) {
                                for (int k = 0; k < nc; k++) pred[k] = ((int64_t)a->val[en * nc + k] + a->val[ep * nc + k]) - a->val[eo * nc + k];
                                ok = 1; a->npar++;
                            }
                        }
                    }
                    if (!ok) for (int k = 0; k < nc; k++) pred[k] = a->val[(p - 1) * nc + k];
                    wrap_add(pred, corr + (size_t)p * nc, nc, a->wmin, a->wmax, a->val + (size_t)p * nc, &a->nwrap);
                }
            } else if (a->pred == 5) { /* TEX_COORDS_PORTABLE, WRAP */
                if (a->xform != 1 || nc != 2 || !posatt) { free(sym); FAIL(UVO_ERR_UNSUPPORTED); }
                int32_t nor = rd_i32(r);
                if (r->err || nor < 0 || nor > n) { free(sym); FAIL(UVO_ERR_CORRUPT); }
                uint8_t *orient = (uint8_t *)malloc((size_t)nor + 1);
                { rabs_t rs; rabs_start(r, &rs); if (r->err) { free(orient); free(sym); FAIL(UVO_ERR_CORRUPT); }
                  int lastv = 1; for (int k = 0; k < nor; k++) { if (!rabs_bit(&rs)) lastv = !lastv; orient[k] = (uint8_t)lastv; } }
                a->wmin = rd_i32(r); a->wmax = rd_i32(r);
                if (r->err || a->wmax < a->wmin) { free(orient); free(sym); FAIL(UVO_ERR_CORRUPT); }
                const int32_t *pos = posatt->val; const int *pv2d = posdec->v2d;
#define POS3(corner, dst) do { int e_ = pv2d[t->c2v[(corner)]]; (dst)[0] = pos[e_ * 3]; (dst)[1] = pos[e_ * 3 + 1]; (dst)[2] = pos[e_ * 3 + 2]; } while (0)
                int bad = 0;
                for (int p = 0; p < n; p++) {
                    int c = d->d2c[p], ncn = c_next(c), pcn = c_prev(c);
                    int nd = d->v2d[v_vert(&vw, ncn)], pd = d->v2d[v_vert(&vw, pcn)];
                    int64_t pred[2]; int have = 0;
                    if (pd < p && nd < p) {
                        int64_t nuv[2] = {a->val[nd * 2], a->val[nd * 2 + 1]}, puv[2] = {a->val[pd * 2], a->val[pd * 2 + 1]};
                        if (nuv[0] == puv[0] && nuv[1] == puv[1]) { pred[0] = puv[0]; pred[1] = puv[1]; have = 1; }
                        else {
                            int64_t tip[3], np3[3], pp3[3]; POS3(c, tip); POS3(ncn, np3); POS3(pcn, pp3);
                            int64_t pn[3] = {pp3[0] - np3[0], pp3[1] - np3[1], pp3[2] - np3[2]};
                            int64_t pn2 = pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2];
                            if (pn2 != 0) {
                                int64_t cnv[3] = {tip[0] - np3[0], tip[1] - np3[1], tip[2] - np3[2]};
                                int64_t dot = pn[0] * cnv[0] + pn[1] * cnv[1] + pn[2] * cnv[2];
                                int64_t pnuv[2] = {puv[0] - nuv[0], puv[1] - nuv[1]};
                                int64_t xuv[2] = {nuv[0] * pn2 + dot * pnuv[0], nuv[1] * pn2 + dot * pnuv[1]};
                                int64_t xpos[3] = {np3[0] + (dot * pn[0]) / pn2, np3[1] + (dot * pn[1]) / pn2, np3[2] + (dot * pn[2]) / pn2};
                                int64_t dx[3] = {tip[0] - xpos[0], tip[1] - xpos[1], tip[2] - xpos[2]};
                                uint64_t cx2 = (uint64_t)(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
                                int64_t ns = (int64_t)int_sqrt(cx2 * (uint64_t)pn2);
                                int64_t cxuv[2] = {pnuv[1] * ns, -pnuv[0] * ns};
                                if (nor <= 0) { bad = 1; break; }
                                int o = orient[--nor];
                                if (o) { pred[0] = (xuv[0] + cxuv[0]) / pn2; pred[1] = (xuv[1] + cxuv[1]) / pn2; }
                                else { pred[0] = (xuv[0] - cxuv[0]) / pn2; pred[1] = (xuv[1] - cxuv[1]) / pn2; }
                                /* upstream narrows the prediction to int before clamping */
                                pred[0] = (int32_t)pred[0]; pred[1] = (int32_t)pred[1];
                                have = 1; a->npar++;
                            }
                        }
                    }
                    if (!have) {
                        if (nd < p) { pred[0] = a->val[nd * 2]; pred[1] = a->val[nd * 2 + 1]; }
                        else if (p > 0) { pred[0] = a->val[(p - 1) * 2]; pred[1] = a->val[(p - 1) * 2 + 1]; }
                        else { pred[0] = pred[1] = 0; }
                    }
                    wrap_add(pred, corr + (size_t)p * 2, 2, a->wmin, a->wmax, a->val + (size_t)p * 2, &a->nwrap);
                }
                a->orient_left = nor;
                free(orient);
                if (bad) { free(sym); FAIL(UVO_ERR_CORRUPT); }
            } else if (a->pred == 6) { /* GEOMETRIC_NORMAL + canonicalised octahedron */
                if (a->xform != 3 || a->seq != 3 || !posatt) { free(sym); FAIL(UVO_ERR_UNSUPPORTED); }
                int32_t max_q = rd_i32(r); int32_t cen_in = rd_i32(r); (void)cen_in;
                rabs_t flips; rabs_start(r, &flips);
                if (r->err || max_q < 3 || (max_q & (max_q + 1)) != 0) { free(sym); FAIL(UVO_ERR_CORRUPT); }
                const int32_t MAXQ = max_q, MAXV = MAXQ - 1, CEN = MAXV / 2;
                a->wmin = MAXQ; a->wmax = CEN;
                const int32_t *pos = posatt->val; const int *pv2d = posdec->v2d;
                for (int p = 0; p < n; p++) {
                    int c0 = d->d2c[p];
                    int64_t cent[3]; POS3(c0, cent);
                    int64_t nrm[3] = {0, 0, 0};
                    int c = c0, left = 1;
                    while (c >= 0) {
                        int64_t pn_[3], pp_[3]; POS3(c_next(c), pn_); POS3(c_prev(c), pp_);
                        int64_t dn[3] = {pn_[0] - cent[0], pn_[1] - cent[1], pn_[2] - cent[2]};
                        int64_t dp[3] = {pp_[0] - cent[0], pp_[1] - cent[1], pp_[2] - cent[2]};
                        nrm[0] += dn[1] * dp[2] - dn[2] * dp[1];
                        nrm[1] += dn[2] * dp[0] - dn[0] * dp[2];
                        nrm[2] += dn[0] * dp[1] - dn[1] * dp[0];
                        int nx;
                        if (left) { nx = v_swl(&vw, c); if (nx < 0) { nx = v_swr(&vw, c0); left = 0; } else if (nx == c0) nx = INV; }
                        else nx = v_swr(&vw, c);
                        c = nx;
                    }
                    int64_t asum = llabs(nrm[0]) + llabs(nrm[1]) + llabs(nrm[2]);
                    if (asum > (1 << 29)) { int64_t q = asum / (1 << 29); nrm[0] /= q; nrm[1] /= q; nrm[2] /= q; }
                    int32_t v[3] = {(int32_t)nrm[0], (int32_t)nrm[1], (int32_t)nrm[2]};
                    int64_t as2 = (int64_t)abs(v[0]) + abs(v[1]) + abs(v[2]);
                    if (as2 == 0) { v[0] = CEN; }
                    else {
                        v[0] = (int32_t)(((int64_t)v[0] * CEN) / as2);
                        v[1] = (int32_t)(((int64_t)v[1] * CEN) / as2);
                        if (v[2] >= 0) v[2] = CEN - abs(v[0]) - abs(v[1]); else v[2] = -(CEN - abs(v[0]) - abs(v[1]));
                    }
                    if (rabs_bit(&flips)) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; a->nflip++; }
                    int32_t s, tt;
                    if (v[0] >= 0) { s = v[1] + CEN; tt = v[2] + CEN; }
                    else { s = v[1] < 0 ? abs(v[2]) : MAXV - abs(v[2]); tt = v[2] < 0 ? abs(v[1]) : MAXV - abs(v[1]); }
                    if ((s == 0 && tt == 0) || (s == 0 && tt == MAXV) || (s == MAXV && tt == 0)) { s = MAXV; tt = MAXV; }
                    else if (s == 0 && tt > CEN) tt = CEN - (tt - CEN);
                    else if (s == MAXV && tt < CEN) tt = CEN + (CEN - tt);
                    else if (tt == MAXV && s < CEN) s = CEN + (CEN - s);
                    else if (tt == 0 && s > CEN) s = CEN - (s - CEN);
                    /* inverse transform: positive corrections */
                    int32_t pr[2] = {s - CEN, tt - CEN};
                    int ind = abs(pr[0]) + abs(pr[1]) <= CEN;
#define INVERT_DIAMOND(px, py) do { int32_t s_ = (px), t_ = (py), ss_, st_; \
    if (s_ >= 0 && t_ >= 0) { ss_ = 1; st_ = 1; } else if (s_ <= 0 && t_ <= 0) { ss_ = -1; st_ = -1; } \
    else { ss_ = s_ > 0 ? 1 : -1; st_ = t_ > 0 ? 1 : -1; } \
    int32_t cs_ = ss_ * CEN, ct_ = st_ * CEN, us_ = 2 * s_ - cs_, ut_ = 2 * t_ - ct_; \
    if (ss_ * st_ >= 0) { int32_t tmp_ = us_; us_ = -ut_; ut_ = -tmp_; } else { int32_t tmp_ = us_; us_ = ut_; ut_ = tmp_; } \
    us_ += cs_; ut_ += ct_; (px) = us_ / 2; (py) = ut_ / 2; } while (0)
#define ROT(px, py, k) do { int32_t x_ = (px), y_ = (py); if ((k) == 1) { (px) = y_; (py) = -x_; } else if ((k) == 2) { (px) = -x_; (py) = -y_; } else if ((k) == 3) { (px) = -y_; (py) = x_; } } while (0)
#define MODMAX(x) ((x) > CEN ? (x) - MAXQ : ((x) < -CEN ? (x) + MAXQ : (x)))
                    if (!ind) INVERT_DIAMOND(pr[0], pr[1]);
                    int bl = (pr[0] == 0 && pr[1] == 0) || (pr[0] < 0 && pr[1] <= 0);
                    int rcnt;
                    if (pr[0] == 0) rcnt = pr[1] == 0 ? 0 : (pr[1] > 0 ? 3 : 1);
                    else if (pr[0] > 0) rcnt = pr[1] >= 0 ? 2 : 1;
                    else rcnt = pr[1] <= 0 ? 0 : 3;
                    if (!bl) ROT(pr[0], pr[1], rcnt);
                    int32_t o0 = pr[0] + corr[p * 2], o1 = pr[1] + corr[p * 2 + 1];
                    o0 = MODMAX(o0); o1 = MODMAX(o1);
                    if (!bl) { int k = (4 - rcnt) % 4; ROT(o0, o1, k); }
                    if (!ind) INVERT_DIAMOND(o0, o1);
                    a->val[p * 2] = o0 + CEN; a->val[p * 2 + 1] = o1 + CEN;
                }
            } else { free(sym); FAIL(UVO_ERR_UNSUPPORTED); }
            free(sym);
            if (r->err) FAIL(UVO_ERR_TRUNCATED);
            if (a->type == 0 && !posatt && a->nc == 3) { posatt = a; posdec = d; }
        }
        for (int j = 0; j < d->natt; j++) {
            attr_t *a = &d->att[j];
            if (a->seq == 2) { for (int k = 0; k < a->nc; k++) a->qmin[k] = rd_f32(r); a->qrange = rd_f32(r); a->qbits = rd_u8(r); }
            else if (a->seq == 3) a->qbits = rd_u8(r);
        }
        if (r->err) FAIL(UVO_ERR_TRUNCATED);
    }
    out->bytes_consumed = r->p;

    /* ---- per-point expansion (DRACOLoader.js:571-590): first attribute of each semantic wins */
    {
        int all_term = 1;
        for (int i = 0; i < ndec; i++) for (int j = 0; j < dec[i].natt; j++) if (!dec[i].att[j].term_ok) all_term = 0;
        out->rans_terminal_ok = all_term;
    }
    int done_sem[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < ndec; i++) {
        adec_t *d = &dec[i];
        const atab_t *av = (d->dec_type == 1) ? &at[d->att_data_id] : NULL;
        for (int j = 0; j < d->natt; j++) {
            attr_t *a = &d->att[j];
            if (a->type < 0 || a->type > 3 || done_sem[a->type]) continue;
            done_sem[a->type] = 1;
            float **dst = a->type == 0 ? &out->position : a->type == 1 ? &out->normal : a->type == 2 ? &out->color : &out->uv;
            int onc = a->nc;
            *dst = (float *)malloc((size_t)np * onc * 4 + 16);
            float delta = 0;
            if (a->seq == 2) { uint32_t maxq = (1u << a->qbits) - 1u; delta = a->qrange / (float)maxq; }
            int32_t MAXV = (a->seq == 3) ? ((1 << a->qbits) - 1) - 1 : 0;
            for (int p = 0; p < np; p++) {
                int c = p2c[p];
                int vert = av ? av->c2v[c] : t->c2v[c];
                int e = d->v2d[vert];
                float *o = *dst + (size_t)p * onc;
                if (a->seq == 2) for (int k = 0; k < onc; k++) o[k] = dequant_value(a->val[e * onc + k], delta, a->qmin[k]);
                else if (a->seq == 3) oct_to_unit(a->val[e * 2], a->val[e * 2 + 1], MAXV, o);
                else {
                    /* INTEGER: ConvertValue<float>; normalized integer types divide by the type max */
                    static const float tmax[12] = {0, 127.f, 255.f, 32767.f, 65535.f, 2147483647.f, 4294967295.f, 0, 0, 0, 0, 0};
                    for (int k = 0; k < onc; k++) {
                        float v = (float)a->val[e * onc + k];
                        if (a->normalized && a->dtype >= 1 && a->dtype <= 6) v = v / tmax[a->dtype];
                        o[k] = v;
                    }
                }
            }
            if (a->type == 0) { out->pos_entries = a->n; out->pos_q = (int32_t *)malloc((size_t)a->n * 3 * 4 + 4); memcpy(out->pos_q, a->val, (size_t)a->n * 3 * 4); out->pos_wraps = a->nwrap; out->pos_parallelograms = a->npar; out->pos_wmin = a->wmin; out->pos_wmax = a->wmax; }
            if (a->type == 3) { out->uv_entries = a->n; out->uv_q = (int32_t *)malloc((size_t)a->n * 2 * 4 + 4); memcpy(out->uv_q, a->val, (size_t)a->n * 2 * 4); out->uv_orient_left = a->orient_left; out->uv_wraps = a->nwrap; out->uv_wmin = a->wmin; out->uv_wmax = a->wmax; }
            if (a->type == 1) { out->nrm_entries = a->n; out->nrm_q = (int32_t *)malloc((size_t)a->n * 2 * 4 + 4); memcpy(out->nrm_q, a->val, (size_t)a->n * 2 * 4); out->nrm_flips = a->nflip; }
        }
    }
    out->dbg_c2v = (int32_t *)malloc((size_t)3 * F * 4); memcpy(out->dbg_c2v, t->c2v, (size_t)3 * F * 4);
    out->dbg_opp = (int32_t *)malloc((size_t)3 * F * 4); memcpy(out->dbg_opp, t->opp, (size_t)3 * F * 4);
done:
    out->status = rc;
    for (int i = 0; i < ndec; i++) { for (int j = 0; j < dec[i].natt; j++) free(dec[i].att[j].val); free(dec[i].d2c); free(dec[i].v2d); }
    for (int i = 0; i < 8; i++) free_atab(&at[i]);
    for (int i = 0; i < 6; i++) free(cn.ctx[i]);
    free(cn.ts); free(cn.t.opp); free(cn.t.c2v); free(cn.t.lmc); free(cn.t.hole);
    free(c2p); free(p2c);
    if (rc) { int s = rc; uvo_draco_free(out); out->status = s; }
    return rc;
#undef FAIL
}
