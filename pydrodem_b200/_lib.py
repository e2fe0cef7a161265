"""ctypes binding of libhydrocore.so (the C ABI in include/hydrocore.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
context is requested, this module raises.  Nothing here imports oracle/.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhydrocore.so")

SEED_WBT, SEED_TAUDEM = 0, 1
TABLE_WBT, TABLE_TAUDEM = 0, 1
DIR_NOFLOW, DIR_NODATA = 0, 255

c_i64, c_i32, c_f32, c_f64, c_vp = (ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_double,
                                    ctypes.c_void_p)

# name -> (restype, argtypes); every symbol include/hydrocore.h declares
SIGNATURES = {
    "hc_last_error": (ctypes.c_char_p, []),
    "hc_version": (ctypes.c_int, []),
    "hc_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int]),
    "hc_destroy": (ctypes.c_int, [c_vp]),
    "hc_workspace_bytes": (c_i64, [c_vp]),
    "hc_launch_count": (c_i64, [c_vp]),
    "hc_fill_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, ctypes.c_int, c_vp]),
    "hc_fill_eps_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, ctypes.c_int, c_f64, ctypes.POINTER(c_i64), c_vp]),
    "hc_d8_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_f64, c_f64, ctypes.c_int, c_vp]),
    "hc_resolve_flats_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, ctypes.c_int, c_vp]),
    "hc_accum_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, c_vp]),
    "hc_fill_d8_accum_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_f64,
                                            c_f64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "hc_fill_d8_accum_ld_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f32, c_f64,
                                               c_f64, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "hc_label_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, ctypes.POINTER(c_i32), c_vp]),
    "hc_fill_single_cell_pits_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_vp]),
    "hc_breach_single_cell_pits_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_vp]),
    "hc_taudem_check_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(c_i64), c_i64, c_i64,
                                           c_f32, c_vp]),
    "hc_diff_round_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, c_vp]),
    "hc_taudem_fix_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_f32, c_vp]),
    "hc_conv2d_symm_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_i64, c_i64, c_vp]),
    "hc_conv2d_symm_dw4_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "hc_slope_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f64, c_f64, ctypes.c_int, c_vp]),
    "hc_slope_thresholds_set": (ctypes.c_int, [c_vp, c_vp, c_i32]),
    "hc_curvature_d2_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f64, c_f64, ctypes.c_int, c_vp]),
    "hc_mean_std_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_vp]),
    "hc_sum_shifted_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, c_f64, ctypes.POINTER(c_f64), c_vp]),
    "hc_breach_flowpath_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_i64, c_f64, c_vp]),
    "hc_wbt_seed_mask_host": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_f32]),
    "hc_breach_flowpath_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_f64, c_f64,
                                              c_i64, c_f64, c_vp, c_vp]),
    "hc_breach_least_cost_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_i64, c_f64, ctypes.c_int, c_f64,
                                                c_vp, c_vp]),
    "hc_band_catalog_local_dev": (ctypes.c_int, [c_vp, c_vp, c_f32] + [c_vp] * 6 + [c_i64, c_i64, ctypes.c_int32, ctypes.c_int32]
                                  + [c_vp] * 13 + [c_i64, ctypes.POINTER(c_i64), c_vp]),
    "hc_band_catalog_owner_ids_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, ctypes.c_int32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hc_catalog_finish_dev": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32, c_vp, c_i64, c_vp, c_vp, ctypes.c_int,
                                             ctypes.c_int32] + [c_vp] * 16 + [c_vp]),
    "hc_band_select_partials_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.c_int32, ctypes.c_int32] + [c_vp] * 6 + [c_vp]),
    "hc_pit_select_tables_dev": (ctypes.c_int, [c_vp, ctypes.c_int32] + [c_vp] * 12 + [c_i64] + [c_vp] * 3
                                 + [c_f64] * 4 + [ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "hc_lzw_decode_host": (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "hc_lzw_encode_host": (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, ctypes.POINTER(c_i64)]),
    "hc_breach_depressions_host": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_f32, c_i64, c_f64, ctypes.c_int,
                                                  ctypes.c_int, c_vp]),
    "hc_minmax_filter_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            c_vp]),
    "hc_conv2d_symm_f64_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_i64, c_i64, c_vp]),
    "hc_river_minimize_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_f32, ctypes.c_int, c_vp, c_i64, c_i64,
                                             c_vp]),
    "hc_raise_small_rivers_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, ctypes.c_int, ctypes.c_int, c_f32,
                                                 c_f32, ctypes.c_int, c_i64, c_i64, c_vp]),
    "hc_raise_small_streams_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_f32, c_f32, ctypes.c_int, c_i64,
                                                  c_i64, c_vp]),
    "hc_fill_missing_data_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, ctypes.c_int, c_f64, c_vp]),
    "hc_mono_mask_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_i64, c_i64, c_vp]),
    "hc_mono_stats_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int32, c_vp, c_i64, c_i64, c_vp]),
    "hc_smooth_rivers_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, ctypes.c_int, c_f32, ctypes.c_int,
                                            c_i64, c_i64, c_vp]),
    "hc_box3_index_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_f32, c_i64, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "hc_lut3d_gather_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(ctypes.c_int),
                                           c_vp]),
    "hc_dtm_select_dev": (ctypes.c_int, [c_vp] + [c_vp] * 13 + [c_i64, c_f32, c_f32, c_vp, c_vp]),
    "hc_dtm_pre_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f64, c_f64, ctypes.c_int, c_i64, c_i64,
                                      ctypes.POINTER(c_f64), c_vp]),
    "hc_dtm_apply_dev": (ctypes.c_int, [c_vp] + [c_vp] * 10 + [c_i64, c_i64, c_f32, c_f32, c_vp, c_vp]),
    "hc_fill_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, ctypes.c_int]),
    "hc_fill_d8_accum_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_f64,
                                             c_f64, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "hc_label_host": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, ctypes.POINTER(c_i32)]),
    "hc_fill_d8_accum_host_begin": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32,
                                                   c_f64, c_f64, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "hc_fill_d8_accum_host_end": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "hc_band_fill_edge_cap": (c_i64, [c_i64]),
    "hc_band_fill_local_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_f32, ctypes.c_int, ctypes.c_int, c_i64, c_vp,
                                              c_vp, c_vp]),
    "hc_band_fill_finish_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "hc_band_d8_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_f64, c_f64, ctypes.c_int,
                                      ctypes.c_int, c_vp]),
    "hc_band_flats_begin_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, ctypes.c_int, ctypes.c_int,
                                               c_vp]),
    "hc_band_flats_get_seams_dev": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "hc_band_flats_put_halos_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.POINTER(ctypes.c_int), c_vp]),
    "hc_band_flats_put_halos_flag_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hc_band_flats_relax_dev": (ctypes.c_int, [c_vp, c_vp]),
    "hc_band_flats_finish_dev": (ctypes.c_int, [c_vp, c_vp]),
    "hc_band_accum_local_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "hc_band_accum_finish_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "hc_band_label_local_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, ctypes.c_int, c_i64, c_vp, c_vp]),
    "hc_band_label_relax_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, ctypes.POINTER(ctypes.c_int), c_vp]),
    "hc_band_label_rank_dev": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), c_vp]),
    "hc_band_label_pairs_dev": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp]),
    "hc_band_label_finish_dev": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "hc_gather_flag_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "hc_band_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]),
    "hc_seg_stats_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hc_pit_catalog_dev": (ctypes.c_int, [c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32,
                                          ctypes.c_int, c_i32] + [c_vp] * 17 + [c_i64, c_vp, ctypes.POINTER(c_i64), c_vp]),
    "hc_pit_select_dev": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32] + [c_vp] * 12 + [c_i64, c_vp, c_vp, c_vp,
                                         c_f64, c_f64, c_f64, c_f64, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "hc_pit_apply_dev": (ctypes.c_int, [c_vp] + [c_vp] * 6 + [c_i64, c_i32, c_i32] + [c_vp] * 6 + [c_i64, c_vp, c_vp, c_vp]),
    "hc_fill_shallow_pits_dev": (ctypes.c_int, [c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32,
                                                ctypes.POINTER(c_i64), c_vp]),
    "hc_kth_value_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, ctypes.POINTER(c_f32), c_vp]),
    "hc_profile_enable": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int]),
    "hc_profile_read": (ctypes.c_int, [c_vp, ctypes.c_char_p, c_i64]),
    "hc_debug_read": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "hc_fill_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i32)]),
}


class HydrocoreError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()


def load():
    """dlopen libhydrocore.so and type every entry point.  Raises if the extension is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise HydrocoreError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C pydrodem_b200/csrc`). pydrodem_b200 has no CPU fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the ABI drifted
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().hc_last_error().decode("utf-8", "replace")
        raise HydrocoreError(f"{what} failed (rc={rc}): {msg}")


_ctxs = {}


def context(device=0):
    """One hc_ctx per device per process (owns the device workspace arena)."""
    lib = load()
    with _lock:
        h = _ctxs.get(device)
        if h is None:
            p = c_vp()
            rc = lib.hc_create(ctypes.byref(p), int(device))
            if rc != 0:
                raise HydrocoreError(f"hc_create(device={device}) failed (rc={rc}): "
                                     f"{lib.hc_last_error().decode('utf-8', 'replace')}")
            h = _ctxs[device] = p
    return h


def new_context(device=0):
    """A private hc_ctx (row bands keep per-band state in their context; several bands can share a
    device).  The caller owns it: free with destroy_context."""
    lib = load()
    p = c_vp()
    rc = lib.hc_create(ctypes.byref(p), int(device))
    if rc != 0:
        raise HydrocoreError(f"hc_create(device={device}) failed (rc={rc}): "
                             f"{lib.hc_last_error().decode('utf-8', 'replace')}")
    return p


def destroy_context(ctx):
    load().hc_destroy(ctx)


def launch_count(device=0, ctx=None):
    return int(load().hc_launch_count(ctx if ctx is not None else context(device)))


def profile_enable(on=True, device=0, max_records=16384, ctx=None):
    check(load().hc_profile_enable(ctx if ctx is not None else context(device), int(bool(on)), int(max_records)),
          "hc_profile_enable")


def profile_read(device=0, ctx=None):
    """{kernel_name: (launches, total_ms, max_ms)} since the last read; synchronises the device."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(load().hc_profile_read(ctx if ctx is not None else context(device), buf, len(buf)), "hc_profile_read")
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, tot, mx = line.split()
        out[name] = (int(cnt), float(tot), float(mx))
    return out


def debug_read(device=0):
    out = (ctypes.c_uint64 * 16)()
    check(load().hc_debug_read(context(device), out), "hc_debug_read")
    return list(out)
