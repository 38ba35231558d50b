"""ctypes binding of the C ABI declared in include/nextpolish_b200.h.

The struct mirrors are the ones the reference's own wrapper declares
(reference source/lib/nextpolish1.py:27-81); the np_* batch functions are this engine's
additions.  Importing this module only loads the shared object: no CUDA call is made until an
engine is created, so it works on machines without a GPU (the compute entry points then fail
loudly — there is no CPU implementation behind them).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.realpath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "nextpolish1.so")


class Configure(C.Structure):  # config.h:25-67
    _fields_ = [
        ("trim_len_edge", C.c_uint8), ("ext_len_edge", C.c_uint8), ("min_map_quality", C.c_uint8),
        ("indel_balance_factor_sgs", C.c_double), ("min_count_ratio_skip", C.c_double),
        ("min_len_ldr", C.c_uint8), ("min_len_inter_kmer", C.c_uint8), ("max_len_kmer", C.c_uint8),
        ("max_count_kmer", C.c_uint8),
        ("min_depth_snp", C.c_uint8), ("min_count_snp", C.c_uint8), ("min_count_snp_link", C.c_int8),
        ("ploidy", C.c_double), ("indel_balance_factor_lgs", C.c_double),
        ("max_indel_factor_lgs", C.c_double), ("max_snp_factor_lgs", C.c_double),
        ("min_snp_factor_sgs", C.c_double),
        ("region_count", C.c_int32), ("count_read_ins_sgs", C.c_uint32), ("max_ins_len_sgs", C.c_uint32),
        ("max_ins_fold_sgs", C.c_int32), ("max_variant_count_lgs", C.c_int32),
        ("max_clip_ratio_sgs", C.c_double), ("max_clip_ratio_lgs", C.c_double),
        ("trace_polish_open", C.c_int32), ("read_tlen", C.c_int32), ("read_len", C.c_int32),
        ("fastafn", C.c_char_p), ("bamfn", C.c_char_p), ("thirdbamfn", C.c_char_p),
    ]


class PolishPoint(C.Structure):  # contig.h:10-15
    _fields_ = [("pos", C.c_int32), ("index", C.c_int16), ("curbase", C.c_char), ("base", C.c_char)]


class PolishResult(C.Structure):  # contig.h:17-22
    _fields_ = [("contig", C.c_void_p), ("data", C.POINTER(PolishPoint)),
                ("length", C.c_int32), ("datalength", C.c_int32)]


class ShardView(C.Structure):  # np_shard_view
    _fields_ = [
        ("n_contigs", C.c_int32), ("n_reads", C.c_int64),
        ("ctg_off", C.POINTER(C.c_int64)), ("ctg_seq", C.c_void_p),
        ("ctg_read_off", C.POINTER(C.c_int64)),
        ("rec_off", C.c_void_p), ("rec", C.c_void_p), ("qual_off", C.c_void_p), ("qual", C.c_void_p),
    ]


class SynthParams(C.Structure):  # np_synth_params
    _fields_ = [
        ("seed", C.c_uint64), ("n_contigs", C.c_int32), ("contig_len", C.c_int64),
        ("min_len", C.c_int64), ("max_len", C.c_int64), ("depth", C.c_double), ("read_len", C.c_int32),
        ("draft_snv", C.c_double), ("draft_indel", C.c_double), ("read_sub", C.c_double),
        ("read_indel", C.c_double), ("lowercase_frac", C.c_double), ("compress_level", C.c_int32),
    ]


class FilesResult(C.Structure):  # np_files_result
    _fields_ = [("task", C.c_int32), ("n_contigs", C.c_int32), ("names", C.POINTER(C.c_char_p)), ("seq", C.c_void_p),
                ("start", C.POINTER(C.c_int64)), ("len", C.POINTER(C.c_int64)),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("load_ms", C.c_float), ("polish_ms", C.c_float)]


EXPORTS = [  # every symbol include/nextpolish_b200.h declares
    "config_init", "config_destory", "score_chain", "kmer_count", "snp_phase", "snp_valid", "lgspolish",
    "polishresult_init", "polishresult_destory",
    "np_shard_load", "np_shard_view_of", "np_shard_contig_name", "np_shard_free", "np_shard_contig_rank",
    "np_shard_algorithmic_bytes",
    "np_engine_create", "np_engine_destroy", "np_last_error", "np_engine_upload", "np_engine_adopt_device",
    "np_engine_run", "np_engine_sync", "np_engine_result_bytes", "np_engine_download",
    "np_engine_result_device", "np_engine_copy_result", "np_engine_pack_result", "np_engine_kernel_times", "np_engine_set_timing", "np_engine_launch_count", "np_engine_window_stats", "np_engine_stream",
    "np_polish_host", "np_synth_write", "np_synth_shard",
    "np_engine_point_count", "np_engine_points",
    "np_bgzf_inflate", "np_shard_load_gpu", "np_shard_load_gpu_seqs", "np_dev_shard_view", "np_dev_shard_contig_name", "np_dev_shard_contig_rank",
    "np_dev_shard_stats", "np_dev_shard_download", "np_dev_shard_free",
    "np_stream_create", "np_stream_destroy", "np_stream_submit", "np_stream_wait", "np_stream_launch_count",
    "np_files_create", "np_files_destroy", "np_files_submit", "np_files_wait",
    "np_resident_create", "np_resident_destroy", "np_resident_submit", "np_resident_wait", "np_resident_launch_count", "np_engine_launch_total",
    "np_multi_create", "np_multi_run", "np_multi_run_names", "np_multi_destroy", "np_partition_contiguous", "np_engine_result_offsets",
    "np_part_plan_create", "np_part_plan_destroy", "np_part_plan_count", "np_part_plan_name", "np_part_plan_finished",
    "np_part_plan_resume_offset", "np_part_record_name", "np_part_open", "np_part_write", "np_part_close",
]


def load(path=None):
    """Load the shared object and declare prototypes. Raises OSError if it is not built."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise OSError("%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.config_init.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    L.config_init.restype = C.POINTER(Configure)
    L.config_destory.argtypes = [C.POINTER(Configure)]
    for fn in ("score_chain", "kmer_count", "snp_phase", "snp_valid", "lgspolish"):
        getattr(L, fn).argtypes = [C.c_char_p, C.POINTER(Configure)]
        getattr(L, fn).restype = C.POINTER(PolishResult)
    L.polishresult_destory.argtypes = [C.POINTER(PolishResult)]
    L.polishresult_init.restype = C.POINTER(PolishResult)
    L.np_last_error.restype = C.c_char_p
    L.np_shard_load.argtypes = [C.c_char_p, C.c_char_p, vp, i32, i32, i32]
    L.np_shard_load.restype = vp
    L.np_shard_view_of.argtypes = [vp, C.POINTER(ShardView)]
    L.np_shard_contig_name.argtypes = [vp, i32]
    L.np_shard_contig_name.restype = C.c_char_p
    L.np_shard_contig_rank.argtypes = [vp, i32]
    L.np_shard_algorithmic_bytes.argtypes = [vp, i32]
    L.np_shard_algorithmic_bytes.restype = i64
    L.np_shard_free.argtypes = [vp]
    L.np_engine_create.argtypes = [i32]
    L.np_engine_create.restype = vp
    L.np_engine_destroy.argtypes = [vp]
    L.np_engine_upload.argtypes = [vp, C.POINTER(ShardView)]
    L.np_engine_adopt_device.argtypes = [vp, C.POINTER(ShardView)]
    L.np_engine_run.argtypes = [vp, i32, C.POINTER(Configure)]
    L.np_engine_sync.argtypes = [vp]
    L.np_engine_result_bytes.argtypes = [vp]
    L.np_engine_result_bytes.restype = i64
    L.np_engine_download.argtypes = [vp, vp, i64, vp]
    L.np_engine_copy_result.argtypes = [vp, vp, i64]
    L.np_engine_pack_result.argtypes = [vp, vp, i64]
    L.np_engine_result_device.argtypes = [vp]
    L.np_engine_result_device.restype = vp
    L.np_engine_kernel_times.argtypes = [vp, vp, vp, i32]
    L.np_engine_set_timing.argtypes = [vp, i32]
    L.np_engine_set_timing.restype = None
    L.np_engine_launch_count.argtypes = [vp]
    L.np_engine_window_stats.argtypes = [vp, vp]
    L.np_engine_stream.argtypes = [vp]
    L.np_engine_stream.restype = vp
    L.np_polish_host.argtypes = [vp, i32, C.POINTER(ShardView), C.POINTER(Configure), vp, i64, vp]
    L.np_bgzf_inflate.argtypes = [i32, vp, i64, vp, i64, vp, vp, vp]
    L.np_engine_point_count.argtypes = [vp]
    L.np_engine_point_count.restype = i64
    L.np_engine_points.argtypes = [vp, vp, i64, vp]
    L.np_shard_load_gpu.argtypes = [i32, C.c_char_p, C.c_char_p, vp, i32, i32]
    L.np_shard_load_gpu.restype = vp
    L.np_dev_shard_view.argtypes = [vp, C.POINTER(ShardView)]
    L.np_dev_shard_view.restype = None
    L.np_dev_shard_contig_name.argtypes = [vp, i32]
    L.np_dev_shard_contig_name.restype = C.c_char_p
    L.np_dev_shard_contig_rank.argtypes = [vp, i32]
    L.np_dev_shard_stats.argtypes = [vp, vp, vp]
    L.np_dev_shard_stats.restype = None
    L.np_dev_shard_download.argtypes = [vp, vp, vp, vp, vp, vp]
    L.np_dev_shard_free.argtypes = [vp]
    L.np_dev_shard_free.restype = None
    L.np_stream_create.argtypes = [i32, i32]
    L.np_stream_create.restype = vp
    L.np_stream_destroy.argtypes = [vp]
    L.np_stream_destroy.restype = None
    L.np_stream_submit.argtypes = [vp, i32, C.POINTER(ShardView), C.POINTER(Configure), vp, i64, vp]
    L.np_stream_submit.restype = i64
    L.np_stream_wait.argtypes = [vp, i64]
    L.np_stream_launch_count.argtypes = [vp]
    L.np_stream_launch_count.restype = i64
    L.np_files_create.argtypes = [i32, i32]
    L.np_files_create.restype = vp
    L.np_files_destroy.argtypes = [vp]
    L.np_files_destroy.restype = None
    L.np_files_submit.argtypes = [vp, i32, C.c_char_p, C.c_char_p, C.POINTER(Configure)]
    L.np_files_submit.restype = i64
    L.np_files_wait.argtypes = [vp, i64, C.POINTER(FilesResult)]
    L.np_resident_create.argtypes = [i32, i32]
    L.np_resident_create.restype = vp
    L.np_resident_destroy.argtypes = [vp]
    L.np_resident_destroy.restype = None
    L.np_resident_submit.restype = i64
    L.np_resident_submit.argtypes = [vp, i32, C.POINTER(ShardView), C.POINTER(Configure), vp, i64]
    L.np_resident_wait.argtypes = [vp, i64, C.POINTER(i64)]
    L.np_resident_launch_count.argtypes = [vp]
    L.np_resident_launch_count.restype = i64
    L.np_engine_launch_total.argtypes = [vp]
    L.np_engine_launch_total.restype = i64
    L.np_multi_create.argtypes = [vp, i32]
    L.np_multi_create.restype = vp
    L.np_multi_run.argtypes = [vp, i32, C.c_char_p, C.c_char_p, C.POINTER(Configure), C.POINTER(FilesResult)]
    L.np_multi_run_names.argtypes = [vp, i32, C.c_char_p, C.c_char_p, C.POINTER(Configure), vp, i32, C.POINTER(FilesResult)]
    L.np_multi_destroy.argtypes = [vp]
    L.np_multi_destroy.restype = None
    L.np_partition_contiguous.argtypes = [vp, i32, i32, vp]
    L.np_partition_contiguous.restype = None
    L.np_engine_result_offsets.argtypes = [vp, vp]
    L.np_part_plan_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    L.np_part_plan_create.restype = vp
    L.np_part_plan_destroy.argtypes = [vp]
    L.np_part_plan_destroy.restype = None
    L.np_part_plan_count.argtypes = [vp]
    L.np_part_plan_name.argtypes = [vp, i32]
    L.np_part_plan_name.restype = C.c_char_p
    L.np_part_plan_finished.argtypes = [vp]
    L.np_part_plan_resume_offset.argtypes = [vp]
    L.np_part_plan_resume_offset.restype = i64
    L.np_part_record_name.argtypes = [C.c_char_p, i32, C.c_char_p, i32]
    L.np_part_open.argtypes = [C.c_char_p, i64]
    L.np_part_open.restype = vp
    L.np_part_write.argtypes = [vp, C.c_char_p, i32, C.c_char_p, i64, i32]
    L.np_part_close.argtypes = [vp]
    L.np_synth_write.argtypes = [C.POINTER(SynthParams), C.c_char_p, C.c_char_p]
    L.np_synth_shard.argtypes = [C.POINTER(SynthParams), i32, i32, i32, i32]
    L.np_synth_shard.restype = vp
    return L
