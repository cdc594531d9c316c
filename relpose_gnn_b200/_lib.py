"""ctypes binding of librpg_b200.so (C ABI: include/rpg.h).

There is NO fallback: if the shared library is missing and cannot be built, importing any op raises.
Structures mirror include/rpg.h field for field; `_check_layout()` compares sizeof/offsetof probes
exported by the library (`rpg_struct_sizes`) against the ctypes mirrors at load time.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librpg_b200.so")

P = C.c_void_p
I = C.c_int
I64 = C.c_int64
U64 = C.c_uint64
F = C.c_float


class RpgError(RuntimeError):
    pass


class Graph(C.Structure):
    _fields_ = [("G", I), ("N", I), ("Ep", I),
                ("src", P), ("dst", P), ("in_ptr", P), ("in_idx", P), ("out_ptr", P), ("out_idx", P),
                ("inv_deg", P), ("deg", P), ("min_ptr", P), ("min_idx", P), ("max_ptr", P), ("max_idx", P),
                ("sel_src", P), ("sel_dst", P), ("sel_patterns", I), ("sel_div", I), ("has_in", P), ("pg_Ep", I), ("pg_N", I)]


class Gemm(C.Structure):
    _fields_ = [("mode", I), ("M", I), ("N", I), ("n_seg", I),
                ("A", P * 6), ("K", I * 6), ("lda", I * 6),
                ("B", P), ("ldb", I), ("R", I), ("splits", I), ("split_stride", I64), ("block_n", I),
                ("bias", P), ("gadd", P * 2), ("gmap", P * 2), ("gadd_ld", I * 2), ("Ep", I), ("Nn", I),
                ("resid", P), ("resid_ld", I), ("row_scale", P), ("row_scale_mod", I),
                ("mask", P), ("mask_ld", I), ("relu", I),
                ("out", P), ("out_relu", P), ("ldo", I), ("out_f32", P), ("ldo_f32", I),
                ("mask_bits", P), ("mask_bits_ld", I), ("out_bits", P), ("out_bits_ld", I),
                ("gadd_f32", P * 2), ("gadd_f32_ld", I * 2), ("resid_lo", P), ("out_lo", P), ("out_relu_lo", P),
                ("n_gseg", I), ("gsel", P * 4), ("gsel_patterns", I), ("gsel_div", I), ("gsrc", P * 4), ("gsrc_ld", I * 4),
                ("gsrc_rows", I), ("a_colsum", P), ("drop_seed", C.c_uint64), ("drop_p", C.c_float)]


class LayerWeights(C.Structure):
    _fields_ = [("D", I)] + [(n, P) for n in (
        "Wn", "W1e_e", "W2e", "W1m_e", "W2m", "Wgtp", "WW", "WWI", "W1u", "W2u",
        "WnT", "W1e_eT", "W2eT", "W1m_eT", "W2mT", "W2uT", "WgtpT", "WWT", "W1uT",
        "b1e", "b2e", "b1m", "b2m", "bgtp", "bW", "b1u", "b2u",
        "Wgc", "WgcT", "WWM", "bgc", "bWm", "Wgtp_f32", "W2m_f32")] + [("variant", I)]


class LayerActs(C.Structure):
    _fields_ = [(n, P) for n in ("x", "e", "P", "h1", "e_new", "e_new_relu", "h2", "m", "gtp", "y", "z", "a",
                                 "h3", "out", "out_relu", "att_aux", "h1_bits", "h2_bits", "h3_bits", "e_new_bits", "out_bits",
                                 "x_bits", "e_bits", "ybar", "mbar")] + [
        ("drop_seed_x", C.c_uint64), ("drop_seed_e", C.c_uint64), ("drop_p", C.c_float), ("gtp16", P)]


class LayerWeightsSplit(C.Structure):
    _fields_ = [("D", I)] + [(n, P) for n in ("Wn3", "W1e_e3", "W2e3", "W1m_e3", "W2m3", "Wgtp3", "WW3", "W1u3", "W2u3",
                                              "b1e", "b2e", "b1m", "b2m", "bgtp", "bW", "b1u", "b2u",
                                              "Wgc3", "WWM3", "bgc", "bWm",
                                              "WnT3", "W1e_eT3", "W2eT3", "W1m_eT3", "W2mT3", "WgcT3", "WWT3", "W1uT3", "W2uT3",
                                              "WnT3_sd", "Wgtp_f32", "W2m_f32")]


class LayerActsSplit(C.Structure):
    _fields_ = [(n, P) for n in ("x_hi", "x_lo", "e_hi", "e_lo", "P", "h1_hi", "h1_lo", "e_new_hi", "e_new_lo",
                                 "e_new_relu_hi", "e_new_relu_lo", "h2_hi", "h2_lo", "m_hi", "m_lo", "gtp", "y_hi", "y_lo",
                                 "z_hi", "z_lo", "a_hi", "a_lo", "h3_hi", "h3_lo", "out_hi", "out_lo", "out_relu_hi",
                                 "out_relu_lo", "ybar_hi", "ybar_lo", "mbar_hi", "mbar_lo", "P_hi", "P_lo",
                                 "h1_bits", "h2_bits", "h3_bits", "e_new_bits", "out_bits", "x_bits", "e_bits")]


_GRAD_NAMES = ("g_mlp0_w", "g_mlp0_b", "g_mlp2_w", "g_mlp2_b", "g_upd0_w", "g_upd0_b", "g_upd2_w", "g_upd2_b",
               "g_edge0_w", "g_edge0_b", "g_edge2_w", "g_edge2_b",
               "g_att_g_w", "g_att_g_b", "g_att_theta_w", "g_att_theta_b", "g_att_phi_w", "g_att_phi_b",
               "g_att_W_w", "g_att_W_b")


class LayerGradsSplit(C.Structure):
    _fields_ = ([(n, P) for n in ("d_out_hi", "d_out_lo", "d_e_new_hi", "d_e_new_lo")] + [("mask_dx", I), ("mask_de", I)] +
                [(n, P) for n in ("dx_hi", "dx_lo", "de_hi", "de_lo", "dh3_hi", "dh3_lo", "dxu_hi", "dxu_lo", "dan_hi", "dan_lo",
                                  "dyn", "dgtp_hi", "dgtp_lo", "Q_hi", "Q_lo", "dh2_hi", "dh2_lo", "de_tot_hi", "de_tot_lo",
                                  "dh1_hi", "dh1_lo", "dP_hi", "dP_lo", "ysum_hi", "ysum_lo", "h2sum_hi", "h2sum_lo",
                                  "split_ws", "colsum_ws", "gtp_bias_tmp", "T_tmp", "Q_f32") + _GRAD_NAMES])


class LayerGrads(C.Structure):
    _fields_ = [("d_out", P), ("d_e_new", P), ("mask_dx", I), ("mask_de", I)] + [(n, P) for n in (
        "dx", "de", "dh3", "dxu", "dan", "dyn", "dgtp", "dm", "dh2", "de_tot", "dh1", "dP", "ysum",
        "split_ws", "colsum_ws", "gtp_bias_tmp", "Q", "h2sum", "T_tmp",
        "g_mlp0_w", "g_mlp0_b", "g_mlp2_w", "g_mlp2_b", "g_upd0_w", "g_upd0_b", "g_upd2_w", "g_upd2_b",
        "g_edge0_w", "g_edge0_b", "g_edge2_w", "g_edge2_b",
        "g_att_g_w", "g_att_g_b", "g_att_theta_w", "g_att_theta_b", "g_att_phi_w", "g_att_phi_b",
        "g_att_W_w", "g_att_W_b")] + [("dxa_ld", I)]


class PackDesc(C.Structure):
    _fields_ = [("src", P), ("dst", P), ("ld_src", C.c_int32), ("r0", C.c_int32), ("c0", C.c_int32), ("rows", C.c_int32),
                ("cols", C.c_int32), ("ld_dst", C.c_int32), ("transpose", C.c_int16), ("dst_f32", C.c_int16),
                ("lo_plane", C.c_int16), ("pad_", C.c_int16)]


PACK_BATCH_MAX = 64


class PackBatch(C.Structure):
    _fields_ = [("d", PackDesc * PACK_BATCH_MAX), ("n", C.c_int32)]


class SgemmDesc(C.Structure):
    _fields_ = [("A", P), ("B", P), ("C", P), ("u", P), ("v", P), ("Cb", P), ("CbT", P)] + [
        (n, C.c_int32) for n in ("M", "N", "K", "lda", "ldb", "ldc", "ldcb", "ldcbT", "transA", "transB", "accumulate", "pad_")]


SGEMM_BATCH_MAX = 8


class SgemmBatch(C.Structure):
    _fields_ = [("d", SgemmDesc * SGEMM_BATCH_MAX), ("n", C.c_int32)]


class ProfRec(C.Structure):
    _fields_ = [("cls", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("ms", C.c_float),
                ("flops", C.c_double), ("bytes", C.c_double)]


PROF_CLASSES = ("gemm_nt", "gemm_tn", "segment_sum", "attention_fwd", "attention_bwd", "edge_init", "reduce", "other")

# name -> (restype, argtypes); every symbol include/rpg.h declares
SIGNATURES = {
    "rpg_last_error_string": (C.c_char_p, []),
    "rpg_version": (I, []),
    "rpg_device_sm_count": (I, [I, C.POINTER(I)]),
    "rpg_launch_count": (I64, []),
    "rpg_profile_begin": (I, []),
    "rpg_profile_end": (I, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(I), C.POINTER(I),
                            C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rpg_profile_records": (I, [C.POINTER(ProfRec), I, C.POINTER(I)]),
    "rpg_template_tables_words": (I64, [I, I]),
    "rpg_template_tables": (I, [P, P, I, I, P, P]),
    "rpg_pack_weights_batch": (I, [C.POINTER(PackBatch), P]),
    "rpg_adam_step": (I, [P, P, P, P, I64, F, F, F, F, F, F, I64, P]),
    "rpg_sgemm_batch": (I, [C.POINTER(SgemmBatch), P]),
    "rpg_validate_edge_index": (I, [P, I64, I, I, I, P, P, P, P]),
    "rpg_selection_patterns": (I, [P, I, I, I, I, P, P]),
    "rpg_gemm": (I, [C.POINTER(Gemm), P]),
    "rpg_set_gemm_cluster": (I, [I]),
    "rpg_wgrad": (I, [P, I, I, P, I, I, I64, P, P, I, P]),
    "rpg_wgrad_bias": (I, [P, I, I, P, I, I, I64, P, P, I, P, P]),
    "rpg_wgrad_blocks": (I, [P, I, I, I, P, I, I, I64, P, P, I, P, P]),
    "rpg_struct_sizes": (None, [C.POINTER(C.c_int32)]),
    "rpg_reduce_splits": (I, [P, I, I64, I, I, P, I, I, P]),
    "rpg_pack_weight": (I, [P, I, I, I, I, I, P, I, I, P]),
    "rpg_cast_f32_to_bf16": (I, [P, P, I64, P]),
    "rpg_cast_bf16_to_f32": (I, [P, P, I64, P]),
    "rpg_attention_series_enabled": (I, []),
    "rpg_attention_fwd_bf16": (I, [P, I64, I, P, I, P]),
    "rpg_attention_bwd_bf16": (I, [P, P, I, C.POINTER(Graph), I64, I, P, I, P]),
    "rpg_attention_fwd": (I, [P, I64, I, P, I, P, P, P]),
    "rpg_cast_f32_to_split": (I, [P, P, P, I64, P]),
    "rpg_split_to_f32": (I, [P, P, P, I64, P]),
    "rpg_pack_weight_lo": (I, [P, I, I, I, I, I, P, I, P]),
    "rpg_edge_init_fwd_f32": (I, [P, I, P, C.POINTER(Graph), I, P, P, I, P]),
    "rpg_aggregate_mean_split": (I, [P, P, I, C.POINTER(Graph), I, P, P, I, P]),
    "rpg_layer_fwd_split": (I, [C.POINTER(LayerWeightsSplit), C.POINTER(Graph), C.POINTER(LayerActsSplit), P]),
    "rpg_layer_bwd_split": (I, [C.POINTER(LayerWeightsSplit), C.POINTER(Graph), C.POINTER(LayerActsSplit),
                                C.POINTER(LayerGradsSplit), P]),
    "rpg_segment_sum_split": (I, [P, P, I, P, P, P, C.POINTER(Graph), I, P, P, I, P]),
    "rpg_attention_bwd_split": (I, [P, P, I, C.POINTER(Graph), I64, I, P, P, I, P]),
    "rpg_wgrad_split": (I, [P, P, I, I, P, P, I, I, I64, P, P, I, P, P]),
    "rpg_edge_init_fwd_split": (I, [P, I, P, C.POINTER(Graph), I, P, P, I, P, P]),
    "rpg_head_bwd_split": (I, [P, P, P, I, I64, I, P, U64, F, P, I, P, P, I, P, P, P, P, I, P, P]),
    "rpg_attention_bwd": (I, [P, P, I, C.POINTER(Graph), I64, I, P, I, P, P]),
    "rpg_aggregate_mean": (I, [P, I, C.POINTER(Graph), I, P, I, P]),
    "rpg_edge_to_node_sum": (I, [P, I, C.POINTER(Graph), I, I, P, I, P]),
    "rpg_segment_sum": (I, [P, I, P, I, P, P, P, C.POINTER(Graph), I, P, I, P]),
    "rpg_segment_sum2": (I, [P, I, C.POINTER(Graph), I, I, I, P, I, P, I, P]),
    "rpg_edge_init_fwd": (I, [P, I, P, C.POINTER(Graph), I, P, I, P, P]),
    "rpg_dropout_mask": (I, [U64, F, I64, I, P, P]),
    "rpg_head_fwd": (I, [P, P, I, I64, I, P, U64, F, P, P, P, P]),
    "rpg_head_bwd_ws_floats": (I64, [I64, I]),
    "rpg_head_bwd": (I, [P, P, I, I64, I, P, U64, F, P, I, P, I, P, P, P, P, I, P, P]),
    "rpg_pose_loss_ws_floats": (I64, [I64]),
    "rpg_pose_loss": (I, [P, P, C.POINTER(Graph), I64, P, P, P, P, P, P]),
    "rpg_pose_criterion": (I, [P, I, P, C.POINTER(Graph), I64, P, P, P, P, P, P, P]),
    "rpg_colsum_bf16": (I, [P, I, I64, I, P, I, P, I, P, P]),
    "rpg_colsum_scratch_floats": (I64, [I64, I]),
    "rpg_layer_fwd": (I, [C.POINTER(LayerWeights), C.POINTER(Graph), C.POINTER(LayerActs), P]),
    "rpg_layer_bwd_ws_floats": (I64, [I, I64, I64]),
    "rpg_reduce_splits_batch": (I, [P, P]),
    "rpg_upload_words": (I, [P, P, I64, P]),
    "rpg_qexp": (I, [P, I64, P, P]),
    "rpg_pose_errors": (I, [P, P, I64, P, P, P]),
    "rpg_knn_graph": (I, [P, I, I, I, I, I, P, P]),
    "rpg_knn_graph_ragged": (I, [P, I, I, P, P, I, I, I, C.c_int64, P, P]),
    "rpg_build_edge_index": (I, [C.POINTER(Graph), P, P]),
    "rpg_per_graph_tables_words": (I64, [I, I, I]),
    "rpg_per_graph_tables": (I, [P, I, I, I, P, P, P]),
    "rpg_selection_patterns_rows": (I, [P, I64, I, I, P, P, P]),
    "rpg_head_bwd_tc_ws_floats": (I64, [I]),
    "rpg_pack_dpose": (I, [P, I64, P, C.c_float, P, P]),
    "rpg_head_bwd_tc": (I, [P, P, I, P, I64, I, C.c_float, P, P, P, I, P, P, P, P, P, P]),
    "rpg_scale_rows": (I, [P, I, I64, I, P, I, P, I, P]),
    "rpg_edge_mask_apply": (I, [P, I64, I, I, P, I, P, P]),
    "rpg_edge_gather": (I, [P, I, I, P, I, I, P, C.POINTER(Graph), I, I, P, P, I, P, P]),
    "rpg_eval_compose": (I, [P, P, C.POINTER(Graph), I, P, P, P, P, P]),
    "rpg_layer_bwd": (I, [C.POINTER(LayerWeights), C.POINTER(Graph), C.POINTER(LayerActs),
                          C.POINTER(LayerGrads), P]),
}

_lib = None
_lock = threading.Lock()


def _check_layout(lib):
    probe = (C.c_int32 * 16)()
    lib.rpg_struct_sizes(probe)
    want = [C.sizeof(Graph), C.sizeof(Gemm), C.sizeof(LayerWeights), C.sizeof(LayerActs), C.sizeof(LayerGrads),
            Gemm.out_f32.offset, LayerGrads.g_mlp0_w.offset, LayerWeights.b1e.offset,
            C.sizeof(LayerWeightsSplit), C.sizeof(LayerActsSplit), C.sizeof(PackDesc), C.sizeof(PackBatch),
            C.sizeof(ProfRec), C.sizeof(SgemmBatch), C.sizeof(LayerGradsSplit), 0]
    if list(probe) != want:
        raise RpgError(f"ctypes mirrors out of sync with include/rpg.h: library {list(probe)} vs python {want}")


def load(build_if_missing=True):
    """Returns the loaded library; raises RpgError if it is missing (no CPU/PyTorch fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH) and build_if_missing:
            from . import build as _build
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise RpgError(f"{LIB_PATH} is missing: build it with `python -m relpose_gnn_b200.build` "
                           "(the CUDA extension is the only implementation; there is no fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _check_layout(lib)
        if os.environ.get("RPG_GEMM_CLUSTER"):          # tuning / A-B knob: 1 = no clusters, 2 = multicast pairs
            if lib.rpg_set_gemm_cluster(int(os.environ["RPG_GEMM_CLUSTER"])) != 0:
                raise RpgError("RPG_GEMM_CLUSTER must be 1 or 2")
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().rpg_last_error_string().decode(errors="replace")
        kind = "argument error" if rc < 0 else "CUDA error"
        raise RpgError(f"{what}: {kind} {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
