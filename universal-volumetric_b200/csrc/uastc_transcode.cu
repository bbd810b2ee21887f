// uastc_transcode.cu -- UASTC 4x4 -> RGBA32 block kernel (placeholder until the restatement lands;
// files are rejected with UVOL_ERR_UNSUPPORTED by the launcher until then).
#include "uvol_ctx.h"
void uvol_uastc_launch(const Ktx2File *, const int32_t *, const uint8_t *, uint8_t *, const uint32_t *, int, uint32_t, cudaStream_t) {}
