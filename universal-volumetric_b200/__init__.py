"""universal-volumetric_b200 -- B200-native UVOL decode hot path (V2 Draco + KTX2/Basis, V1 Corto).

The product is ``libuvol_b200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/uvol_b200.h``); this package is the thin host-side mirror of the reference's loader
interface.  Import it with ``importlib.import_module("universal-volumetric_b200")``.
"""
from . import _native
from ._native import UvolError, MEM_DEVICE, MEM_HOST, TEX_ASTC_4x4, TEX_BC1, TEX_BC3, TEX_BC7, TEX_ETC1, TEX_ETC2_RGBA, TEX_RGBA32
from .loaders import Context, CortoDecoder, DRACOLoader, KTX2Loader, V2Player, ktx2_probe, pick_texture_format, span_ms
from .manifest import V1Manifest, V1Sequence, V2Manifest, V2Playback, V2Sequence, check_total_frames, emit_v1, emit_v2, normalize_v2, shard_v2
from . import gather

__all__ = ["Context", "DRACOLoader", "KTX2Loader", "V2Player", "CortoDecoder", "V1Manifest", "V1Sequence", "V2Manifest", "V2Sequence", "V2Playback", "shard_v2", "emit_v2", "emit_v1", "normalize_v2", "check_total_frames", "gather", "span_ms", "ktx2_probe", "pick_texture_format", "UvolError", "MEM_DEVICE", "MEM_HOST", "TEX_RGBA32", "TEX_ETC1", "TEX_BC7", "TEX_ASTC_4x4", "TEX_ETC2_RGBA", "TEX_BC1", "TEX_BC3", "_native"]
