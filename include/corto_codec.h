/*
 * corto_codec.h -- the reference's own V1 C ABI, exported unchanged by libuvol_b200.so.
 *
 * Replaces deprecated/encoder/dev/src/corto_codec.h:17-44 (implementation corto_codec.cpp:6-59), the
 * interface the Unity player binds with P/Invoke (deprecated/unity/Assets/uvol/unity/CortoMeshLoader.cs:13-28,63-68).
 * Same three symbols, same struct layouts, same ownership: the caller allocates every output array from
 * decoderInfo[0] = {x: nface, y: nvert} (CortoMeshLoader.cs:16-20); the library owns only the opaque handle.
 * Differences: decoding runs on CUDA device 0 (no CPU path: DecodeMesh returns UVOL_STATUS_CUDA = -4 without
 * a device); errors are negative return values instead of C++ exceptions escaping the ABI; the input needs no
 * 4-byte alignment (decoder.cpp:42-43) because CreateDecoder copies it; normals / colours are not decoded
 * (UVOL V1 frames carry position + uv only, src/V1/player.ts:292-294) and those arrays are left untouched.
 */
#ifndef UVOL_CORTO_CODEC_H
#define UVOL_CORTO_CODEC_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct Color { float r, g, b, a; } Color;
typedef struct Vector2 { float x, y; } Vector2;
typedef struct Vector3 { float x, y, z; } Vector3;
typedef struct Decoder Decoder;

Decoder *CreateDecoder(int length, unsigned char *data, Vector2 *decoderInfo);
void DestroyDecoder(Decoder *decoder);
/* returns nface, -1 for point clouds (corto_codec.cpp:27-30), or a negative uvol_status */
int DecodeMesh(Decoder *decoder, Vector3 *vertices, int *indices, Vector3 *normals, Color *colors, Vector2 *texcoord);

#ifdef __cplusplus
}
#endif
#endif
