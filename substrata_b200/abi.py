"""ctypes declarations of include/sbm_audio.h (kept in one place; tests/test_abi.py checks them against the header)."""
import ctypes as C

_vp, _i, _f, _sz, _u64, _i64, _d = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_uint64, C.c_int64, C.c_double

ERROR_NAMES = {0: "SBM_OK", -1: "SBM_ERR_INVALID_ARG", -2: "SBM_ERR_CUDA", -3: "SBM_ERR_NO_MEMORY", -4: "SBM_ERR_UNKNOWN_ID",
               -5: "SBM_ERR_CAPACITY", -6: "SBM_ERR_UNSUPPORTED", -7: "SBM_ERR_PEER_TIMEOUT"}


class SbmConfig(C.Structure):
    _fields_ = [("frames_per_buffer", _i), ("sample_rate_hz", _i), ("device", _i), ("max_sources", _i), ("max_mix_sources", _i),
                ("sh_hrir", C.POINTER(C.c_float)), ("sh_hrir_taps", _i)]


class SbmSourceState(C.Structure):
    _fields_ = [("cur_read_i", _u64), ("eof_marker", _i64), ("ring_size", _u64), ("mix_pos0", _d), ("direct_gain", _f),
                ("occlusion", _f), ("alive", _i)]


class SbmStats(C.Structure):
    _fields_ = [("buffers_mixed", _u64), ("kernel_launches", _u64), ("active_sources", C.c_uint32), ("last_params_kernel_ms", _f),
                ("last_mix_kernel_ms", _f), ("last_finish_kernel_ms", _f), ("last_algorithmic_bytes", _u64), ("last_h2d_bytes", _u64),
                ("last_d2h_bytes", _u64), ("last_rendered_sources", C.c_uint32), ("ignored_setter_items", C.c_uint32)]


# name -> (restype, argtypes); every SBM_API symbol of the header.
SIGNATURES = {
    "sbm_create": (_i, [_vp, _vp]),
    "sbm_destroy": (None, [_vp]),
    "sbm_last_error": (C.c_char_p, [_vp]),
    "sbm_buffer_create": (_i, [_vp, _vp, _sz, _vp]),
    "sbm_buffer_adopt_device": (_i, [_vp, _vp, _sz, _sz, _i, _vp]),
    "sbm_buffer_destroy": (_i, [_vp, _i]),
    "sbm_source_create": (_i, [_vp, _i, _i, _vp]),
    "sbm_source_destroy": (_i, [_vp, _i, _vp]),
    "sbm_source_attach_buffer": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sbm_source_attach_ring": (_i, [_vp, _i, _i, _i]),
    "sbm_source_ring_push": (_i, [_vp, _i, _vp, _sz]),
    "sbm_source_ring_push_batch": (_i, [_vp, _i, _vp, _vp, _vp]),
    "sbm_host_alloc_pinned": (_vp, [_sz]),
    "sbm_host_free_pinned": (None, [_vp]),
    "sbm_source_ring_size": (_i, [_vp, _i, _vp]),
    "sbm_source_attach_mix": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sbm_source_attach_mix1": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sbm_source_set_mix_params": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "sbm_source_set_sampling_rate": (_i, [_vp, _i, _i]),
    "sbm_source_set_eof_marker": (_i, [_vp, _i, _i64]),
    "sbm_source_set_looping": (_i, [_vp, _i, _i]),
    "sbm_set_source_positions": (_i, [_vp, _i, _vp, _vp]),
    "sbm_set_source_volumes": (_i, [_vp, _i, _vp, _vp]),
    "sbm_set_source_occlusions": (_i, [_vp, _i, _vp, _vp]),
    "sbm_set_head": (_i, [_vp, _vp, _vp]),
    "sbm_set_master_volume": (_i, [_vp, _f]),
    "sbm_set_room_sh_hrir": (_i, [_vp, _vp, _i]),
    "sbm_set_reflection_properties": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f]),
    "sbm_enable_room_effects": (_i, [_vp, _i]),
    "sbm_mix": (_i, [_vp, _vp, _vp]),
    "sbm_mix_submit": (_i, [_vp]),
    "sbm_mix_collect": (_i, [_vp, _vp, _vp]),
    "sbm_mix_begin": (_i, [_vp]),
    "sbm_mix_finish": (_i, [_vp, _vp, _vp]),
    "sbm_peer_export": (_i, [_vp, _vp, _sz]),
    "sbm_set_peer_timeout_ms": (_i, [_vp, _i]),
    "sbm_peer_resync": (_i, [_vp]),
    "sbm_peer_connect": (_i, [_vp, _i, _i, _vp, _sz]),
    "sbm_bus_device": (_i, [_vp, _vp, _vp]),
    "sbm_out_device": (_i, [_vp, _vp]),
    "sbm_stream": (_vp, [_vp]),
    "sbm_poll_finished": (_i, [_vp, _vp, _i, _vp]),
    "sbm_source_query": (_i, [_vp, _i, _vp]),
    "sbm_get_stats": (_i, [_vp, _vp]),
    "sbm_set_profiling": (_i, [_vp, _i]),
    "sbm_resampler_create": (_i, [_i, _i, _vp, _i, _vp]),
    "sbm_resampler_destroy": (None, [_vp]),
    "sbm_resampler_num_src_needed": (_i, [_vp, _sz, _vp]),
    "sbm_resampler_resample": (_i, [_vp, _vp, _sz, _vp, _vp, _sz]),
    "sbm_polyphase_create": (_i, [_i, _vp]),
    "sbm_polyphase_destroy": (None, [_vp]),
    "sbm_polyphase_rates_supported": (_i, [_i, _i]),
    "sbm_polyphase_set_rate_and_channels": (_i, [_vp, _i, _i, _i]),
    "sbm_polyphase_reset_state": (None, [_vp]),
    "sbm_polyphase_max_output_length": (_sz, [_vp, _sz]),
    "sbm_polyphase_next_output_length": (_sz, [_vp, _sz]),
    "sbm_polyphase_process": (_i, [_vp, _vp, _sz, _vp, _sz]),
    "sbm_resample_sh_hrir": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "sbm_debug_ramp_advance": (_f, [_f, _f, C.c_uint32]),
    "sbm_debug_atan2f": (_f, [_f, _f]),
    "sbm_debug_encoder_table": (_vp, []),
    "sbm_debug_get_ambisonic_bus": (_i, [_vp, _vp]),
}


def bind(cdll, require_all=False):
    """Sets restype/argtypes on every symbol the library exports.  The host-emulation test build lacks the resampler."""
    for name, (res, args) in SIGNATURES.items():
        try:
            f = getattr(cdll, name)
        except AttributeError:
            if require_all:
                raise
            continue
        f.restype, f.argtypes = res, args
    return cdll
