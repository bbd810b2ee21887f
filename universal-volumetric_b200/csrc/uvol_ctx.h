// uvol_ctx.h -- per-GPU context: streams, grow-only device / pinned-host arenas, error text.
// One ctx per GPU; calls on a ctx are stream-ordered (mirrors "one decoder instance per worker",
// src/lib/DRACOLoader.js:439).  No torch types, no CPU decode path: every entry point fails with
// UVOL_ERR_CUDA when the device is unusable.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <mutex>
#include <stdlib.h>
#include <thread>
#include <string>
#include <vector>
#include "uvol_internal.h"
#include "../../include/uvol_b200.h"

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + (n / 8 < (256u << 20) ? n / 8 : (256u << 20)) + (1 << 20);      // grow-only with bounded slack
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + (n / 8 < (256u << 20) ? n / 8 : (256u << 20)) + (1 << 20);
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct GeoBatch; struct TexBatch; struct CortoBatch;
void uvol_geo_batch_free(GeoBatch *); void uvol_tex_batch_free(TexBatch *); void uvol_corto_batch_free(CortoBatch *);

struct uvol_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t ev[32] = {}, aux_ev[8] = {}, tex_ev[8] = {}, sync_ev[12] = {};
    cudaStream_t s2 = nullptr;
    cudaStream_t s3 = nullptr;                    // early result copies (index buffers) next to the geometry kernels
    cudaStream_t s4 = nullptr;                    // texture result copies of the pipelined fresh path (chunk k's copy runs next to chunk k+1's upload and kernels)
    cudaEvent_t tex_chunk_ev[16] = {};
    std::string err; std::mutex err_mu;           // the texture side of uvol_decode_v2_batch runs on a helper thread: error text is set under err_mu
    void set_error(const char *msg) { std::lock_guard<std::mutex> g(err_mu); err = msg; }
    uvol_config cfg = {};                          // defaults + environment overrides (uvol_config_default), or the caller's (uvol_create_with_config)
    // geometry path
    PinBuf h_blob, h_desc, h_aux, h_counts, h_out;
    DevBuf d_blob, d_desc, d_aux, d_counts, d_scratch, d_zscratch, d_scratch2, d_zscratch2, d_jobs;
    DevBuf d_out_geo;             // library-owned geometry outputs (valid until the next geometry batch on this ctx)
    // texture path
    PinBuf h_tblob, h_tdesc, h_tout, h_tstate;
    DevBuf d_tblob, d_tdesc, d_tslices, d_tscratch, d_out_tex;
    // V1 path
    PinBuf h_cblob, h_cdesc, h_cout, h_ccounts;
    DevBuf d_cblob, d_cdesc, d_cscratch, d_czscratch, d_out_corto, d_ccounts, d_caux;
    GeoBatch *geo = nullptr; TexBatch *tex = nullptr; CortoBatch *corto = nullptr;
    DevBuf d_flush;
    // stats of the last batch
    uvol_stats stats = {}, stats_tex = {};
    PinBuf *ph_out = &h_out, *ph_tout = &h_tout;   // pinned host outputs: own, or another ctx's (uvol_share_host_outputs)
    bool profile = false;
    int span_geo_end = 0, span_tex_end = 0;      // event index of the last kernel stamp of the last geometry / texture run (0: none)
};

#define UVOL_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    char b_[256]; snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); (ctx)->set_error(b_); return UVOL_ERR_CUDA; } } while (0)

// Host threads used to stage a large batch into the pinned input blob: UVOL_STAGING_THREADS, else min(8, cores / 2).  A launcher that
// runs one process per GPU divides the cores between its ranks through the variable.
static inline int uvol_staging_threads() {
    static const int n = [] { const char *e = getenv("UVOL_STAGING_THREADS"); int v = e ? atoi(e) : 0;
                              if (v <= 0) { v = (int)std::thread::hardware_concurrency() / 2; if (v > 8) v = 8; }
                              return v < 1 ? 1 : (v > 64 ? 64 : v); }();
    return n;
}

static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

int uvol_draco_parse(const uint8_t *data, size_t len, DracoFrame &f, std::vector<uint32_t> &aux);
