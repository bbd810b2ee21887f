// uvol_ctx.h -- per-GPU context: streams, grow-only device / pinned-host arenas, error text.
// One ctx per GPU; calls on a ctx are stream-ordered (mirrors "one decoder instance per worker",
// src/lib/DRACOLoader.js:439).  No torch types, no CPU decode path: every entry point fails with
// UVOL_ERR_CUDA when the device is unusable.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
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
        size_t want = n + n / 8 + (1 << 20);
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
        size_t want = n + n / 8 + (1 << 20);
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// The large arenas of the V2 path (scratch, device outputs, pinned host outputs).  A ctx normally owns its set; contexts
// created for consecutive WINDOWS of one sequence can share one set (uvol_share_arenas), so each window keeps its
// compressed inputs resident while the scratch exists once.
struct Arenas {
    DevBuf d_scratch, d_zscratch, d_scratch2, d_zscratch2, d_out_geo, d_tscratch, d_out_tex;
    PinBuf h_out, h_tout;
};

struct GeoBatch; struct TexBatch; struct CortoBatch;
void uvol_geo_batch_free(GeoBatch *); void uvol_tex_batch_free(TexBatch *); void uvol_corto_batch_free(CortoBatch *);

struct uvol_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t ev[32] = {}, aux_ev[8] = {}, tex_ev[8] = {}, sync_ev[8] = {};
    cudaStream_t s2 = nullptr;
    std::string err;
    // geometry path
    Arenas own_arenas; Arenas *ar = &own_arenas;      // scratch + library-owned outputs (valid until the next batch on any ctx sharing them)
    PinBuf h_blob, h_desc, h_aux, h_counts;
    DevBuf d_blob, d_desc, d_aux, d_counts, d_jobs;
    // texture path
    PinBuf h_tblob, h_tdesc;
    DevBuf d_tblob, d_tdesc, d_tslices;
    // V1 path
    PinBuf h_cblob, h_cdesc, h_cout, h_ccounts;
    DevBuf d_cblob, d_cdesc, d_cscratch, d_czscratch, d_out_corto, d_ccounts, d_caux;
    GeoBatch *geo = nullptr; TexBatch *tex = nullptr; CortoBatch *corto = nullptr;
    DevBuf d_flush;
    // stats of the last batch
    uvol_stats stats = {}, stats_tex = {};
    bool profile = false;
};

#define UVOL_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    char b_[256]; snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); (ctx)->err = b_; return UVOL_ERR_CUDA; } } while (0)

static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

int uvol_draco_parse(const uint8_t *data, size_t len, DracoFrame &f, std::vector<uint32_t> &aux);
