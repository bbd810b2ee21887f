// uvol_api.cu -- context management and the small C-ABI utilities of libuvol_b200.so.
#include <string.h>
#include <stdlib.h>
#include "uvol_ctx.h"

extern "C" const char *uvol_geo_stage_name(int i);
extern "C" const char *uvol_tex_stage_name(int i);
extern "C" const char *uvol_corto_stage_name(int i);

// Defaults of every tunable, then the environment overrides (SURVEY 5 "config / flags": the reference takes constructor arguments
// only -- bufferDuration / intervalDuration, src/Player.ts:50-51 -- and picks the texture target from the GPU, KTX2Loader.js:591-689).
extern "C" void uvol_config_default(uvol_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = (uint32_t)sizeof *cfg;
    cfg->texture_target = UVOL_TEX_RGBA32; cfg->corto_index_u16 = 0; cfg->staging_threads = 0;
    cfg->max_faces_per_frame = 1ull << 24; cfg->max_texture_bytes = 1ull << 31;
    cfg->buffer_duration_s = 4.0; cfg->interval_duration_s = 2.0;
    auto env_u64 = [](const char *name, uint64_t &v) { const char *e = getenv(name); if (e && *e) { char *end = nullptr; const unsigned long long x = strtoull(e, &end, 0); if (end && *end == 0) v = x; } };
    auto env_f64 = [](const char *name, double &v) { const char *e = getenv(name); if (e && *e) { char *end = nullptr; const double x = strtod(e, &end); if (end && *end == 0 && x > 0) v = x; } };
    uint64_t t = cfg->texture_target, u16 = 0, thr = 0;
    if (const char *e = getenv("UVOL_TEXTURE_TARGET")) {
        if (!strcmp(e, "rgba32")) t = UVOL_TEX_RGBA32; else if (!strcmp(e, "etc1")) t = UVOL_TEX_ETC1; else if (!strcmp(e, "bc7")) t = UVOL_TEX_BC7; else if (!strcmp(e, "astc")) t = UVOL_TEX_ASTC_4x4; else if (!strcmp(e, "etc2")) t = UVOL_TEX_ETC2_RGBA; else if (!strcmp(e, "bc1")) t = UVOL_TEX_BC1; else if (!strcmp(e, "bc3")) t = UVOL_TEX_BC3; else env_u64("UVOL_TEXTURE_TARGET", t);
    }
    env_u64("UVOL_CORTO_INDEX_U16", u16); env_u64("UVOL_STAGING_THREADS", thr);
    cfg->texture_target = (uint32_t)t; cfg->corto_index_u16 = (uint32_t)(u16 != 0); cfg->staging_threads = (uint32_t)thr;
    env_u64("UVOL_MAX_FACES", cfg->max_faces_per_frame); env_u64("UVOL_MAX_TEXTURE_BYTES", cfg->max_texture_bytes);
    env_f64("UVOL_BUFFER_DURATION", cfg->buffer_duration_s); env_f64("UVOL_INTERVAL_DURATION", cfg->interval_duration_s);
}

extern "C" int uvol_create(int device, uvol_ctx **out) { return uvol_create_with_config(device, nullptr, out); }

extern "C" int uvol_get_config(const uvol_ctx *c, uvol_config *out) { if (!c || !out) return UVOL_ERR_ARG; *out = c->cfg; return UVOL_OK; }

extern "C" int uvol_create_with_config(int device, const uvol_config *cfg, uvol_ctx **out) {
    if (!out) return UVOL_ERR_ARG;
    if (cfg && cfg->struct_size != sizeof(uvol_config)) return UVOL_ERR_ARG;
    if (cfg && (cfg->texture_target > UVOL_TEX_BC3 || cfg->texture_target == UVOL_TEX_ETC2_RGB)) return UVOL_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return UVOL_ERR_CUDA;   // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return UVOL_ERR_CUDA;
    uvol_ctx *c = new uvol_ctx();
    c->device = device;
    if (cfg) c->cfg = *cfg; else uvol_config_default(&c->cfg);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    // s0 carries the geometry critical path and gets the highest priority: the block scheduler serves pending grids in
    // priority order, so the wide side-stream grids (entropy runs on s1, textures on s2) never hold back a main-stream kernel.
    int prio_lo = 0, prio_hi = 0; cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const int prio_mid = prio_hi < prio_lo - 1 ? prio_hi + 1 : prio_lo;
    if (cudaStreamCreateWithPriority(&c->s0, cudaStreamNonBlocking, prio_hi) != cudaSuccess || cudaStreamCreateWithPriority(&c->s1, cudaStreamNonBlocking, prio_mid) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    if (cudaStreamCreateWithPriority(&c->s2, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    if (cudaStreamCreateWithPriority(&c->s3, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    if (cudaStreamCreateWithPriority(&c->s4, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    for (auto &e : c->tex_chunk_ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    for (auto &e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    for (auto &e : c->aux_ev) if (cudaEventCreate(&e) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    for (auto &e : c->tex_ev) if (cudaEventCreate(&e) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    for (auto &e : c->sync_ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { delete c; return UVOL_ERR_CUDA; }
    *out = c;
    return UVOL_OK;
}

extern "C" void uvol_destroy(uvol_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf *db[] = {&c->d_blob, &c->d_desc, &c->d_aux, &c->d_counts, &c->d_scratch, &c->d_zscratch, &c->d_scratch2, &c->d_zscratch2, &c->d_jobs, &c->d_out_geo,
                    &c->d_tblob, &c->d_tdesc, &c->d_tslices, &c->d_tscratch, &c->d_out_tex,
                    &c->d_cblob, &c->d_cdesc, &c->d_cscratch, &c->d_czscratch, &c->d_out_corto, &c->d_ccounts, &c->d_caux};
    for (auto *b : db) b->release();
    PinBuf *pb[] = {&c->h_blob, &c->h_desc, &c->h_aux, &c->h_counts, &c->h_out, &c->h_tblob, &c->h_tdesc, &c->h_tout, &c->h_tstate, &c->h_cblob, &c->h_cdesc, &c->h_cout, &c->h_ccounts};
    for (auto *b : pb) b->release();
    c->d_flush.release();
    if (c->geo) uvol_geo_batch_free(c->geo);
    if (c->tex) uvol_tex_batch_free(c->tex);
    if (c->corto) uvol_corto_batch_free(c->corto);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->aux_ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->tex_ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->sync_ev) if (e) cudaEventDestroy(e);
    if (c->s2) cudaStreamDestroy(c->s2);
    if (c->s3) cudaStreamDestroy(c->s3);
    if (c->s4) cudaStreamDestroy(c->s4);
    for (auto &e : c->tex_chunk_ev) if (e) cudaEventDestroy(e);
    if (c->s0) cudaStreamDestroy(c->s0);
    if (c->s1) cudaStreamDestroy(c->s1);
    delete c;
}

// Kept for ABI compatibility: since the traversal records shrank to 32 bytes per face and alias the connectivity temporaries, a
// sequence's scratch fits one context and windows no longer need a shared arena.  No effect.
extern "C" int uvol_share_arenas(uvol_ctx *ctx, uvol_ctx *owner) {
    if (!ctx || !owner || ctx->device != owner->device) return UVOL_ERR_ARG;
    return UVOL_OK;
}

// Lets `ctx` write its UVOL_MEM_HOST results into the pinned host buffers of `owner` instead of its own (bounds the pinned memory
// of a windowed sequence to one window).  The results of either ctx are then valid only until the next UVOL_MEM_HOST call on the
// other, and such calls must not run concurrently.
extern "C" int uvol_share_host_outputs(uvol_ctx *ctx, uvol_ctx *owner) {
    if (!ctx || !owner || ctx->device != owner->device) return UVOL_ERR_ARG;
    ctx->ph_out = owner->ph_out; ctx->ph_tout = owner->ph_tout;
    return UVOL_OK;
}

// Device time spanned by the last V2 / geometry / texture calls of `n` contexts that ran concurrently: from the first kernel
// of any of them to the last kernel of any of them (CUDA events of the contexts' streams; requires profiling on).
extern "C" int uvol_span_ms(uvol_ctx *const *ctxs, int n, float *ms) {
    if (!ctxs || n <= 0 || !ms || !ctxs[0]) return UVOL_ERR_ARG;
    float lo = 0.f, hi = 0.f; bool any = false;
    cudaEvent_t ref = nullptr;
    for (int i = 0; i < n && !ref; i++) if (ctxs[i] && ctxs[i]->span_geo_end > 0) ref = ctxs[i]->ev[1];
    for (int i = 0; i < n && !ref; i++) if (ctxs[i] && ctxs[i]->span_tex_end > 0) ref = ctxs[i]->tex_ev[1];
    if (!ref) return UVOL_ERR_ARG;
    auto add = [&](cudaEvent_t b, cudaEvent_t e) {
        float tb = 0.f, te = 0.f;
        if (cudaEventElapsedTime(&tb, ref, b) != cudaSuccess || cudaEventElapsedTime(&te, ref, e) != cudaSuccess) { cudaGetLastError(); return; }
        if (!any || tb < lo) lo = tb;
        if (!any || te > hi) hi = te;
        any = true;
    };
    for (int i = 0; i < n; i++) {
        uvol_ctx *c = ctxs[i]; if (!c) continue;
        if (c->span_geo_end > 0) add(c->ev[1], c->ev[c->span_geo_end]);
        if (c->span_tex_end > 0) add(c->tex_ev[1], c->tex_ev[c->span_tex_end]);
    }
    if (!any) return UVOL_ERR_ARG;
    *ms = hi - lo;
    return UVOL_OK;
}

extern "C" int uvol_release(uvol_ctx *c) {
    if (!c) return UVOL_ERR_ARG;
    UVOL_CUDA(c, cudaSetDevice(c->device));
    UVOL_CUDA(c, cudaDeviceSynchronize());
    c->d_out_geo.release(); c->d_out_tex.release(); c->d_out_corto.release();
    c->h_out.release(); c->h_tout.release(); c->h_cout.release();
    return UVOL_OK;
}

extern "C" const char *uvol_last_error(const uvol_ctx *c) { return c ? c->err.c_str() : "null ctx"; }
extern "C" int uvol_get_stats(const uvol_ctx *c, uvol_stats *out) { if (!c || !out) return UVOL_ERR_ARG; *out = c->stats; return UVOL_OK; }
extern "C" int uvol_set_profiling(uvol_ctx *c, int enable) { if (!c) return UVOL_ERR_ARG; c->profile = enable != 0; return UVOL_OK; }
extern "C" const char *uvol_stage_name(int kind, int stage) {
    if (kind == 0) return uvol_geo_stage_name(stage);
    if (kind == 1) return uvol_tex_stage_name(stage);
    if (kind == 2) return uvol_corto_stage_name(stage);
    return "";
}

extern "C" int uvol_flush_l2(uvol_ctx *c) {
    if (!c) return UVOL_ERR_ARG;
    UVOL_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = 256ull << 20;
    UVOL_CUDA(c, c->d_flush.reserve(bytes));
    UVOL_CUDA(c, cudaMemsetAsync(c->d_flush.p, 0x5a, bytes, c->s0));
    UVOL_CUDA(c, cudaStreamSynchronize(c->s0));
    return UVOL_OK;
}
