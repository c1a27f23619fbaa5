"""alive_vc_b200 - B200-native kNN voice-library matching (ALiVE-VC hot path).

Public surface = the reference's own Python symbols for this path:
    match_features(source, reference, k=4, alpha=0.0)      module/common.py:96
    VoiceLibrary(num_tokens=512, hubert_dim=768)           module/voice_library.py:6
plus the packed-library handles the kernels work on (pack_library, match_packed,
match_indices, ShardedLibrary).  Everything runs through the C ABI in
include/alive_knn.h (alive_vc_b200/libalive_knn.so, sm_100a); there is no fallback.
"""
from .matching import (PackedFrames, StreamingMatcher, clear_pack_cache, load_packed_library, save_packed_library, match_features, match_indices, match_packed,
                       match_packed_queries, pack_frames, pack_libraries, pack_library, pack_queries, search_topk)
from .voice_library import VoiceLibrary
from .lifecycle import HostPipeline, HostStreamingMatcher, LibraryBuilder, RowsContentEncoder, match_rows, match_windows, pack_rows

__all__ = [
    "match_features", "VoiceLibrary", "match_indices", "match_packed", "pack_library", "pack_libraries", "pack_frames",
    "search_topk", "PackedFrames", "clear_pack_cache", "StreamingMatcher",
    "save_packed_library", "load_packed_library", "LibraryBuilder", "match_windows", "HostStreamingMatcher",
    "match_rows", "pack_rows", "match_packed_queries", "pack_queries", "RowsContentEncoder", "HostPipeline",
]
