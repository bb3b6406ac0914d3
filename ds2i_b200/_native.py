"""ctypes binding of libds2i_gpu.so (include/ds2i_gpu.h).  No fallback: if the library is missing
or CUDA is unavailable every call raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DS2I_GPU_LIB selects another build of the same library (kernel experiments: tools/kbench.py)
LIB_PATH = os.environ.get("DS2I_GPU_LIB") or os.path.join(_HERE, "lib", "libds2i_gpu.so")

# every symbol include/ds2i_gpu.h declares
SYMBOLS = [
    "ds2i_gpu_last_error", "ds2i_gpu_op_from_name",
    "ds2i_gpu_index_open", "ds2i_gpu_index_open_file", "ds2i_gpu_index_close", "ds2i_gpu_index_size",
    "ds2i_gpu_index_num_docs", "ds2i_gpu_index_device_bytes", "ds2i_gpu_index_set_global_stats", "ds2i_gpu_index_list_sizes",
    "ds2i_gpu_wand_open", "ds2i_gpu_wand_open_file", "ds2i_gpu_wand_close",
    "ds2i_gpu_query_batch", "ds2i_gpu_query_batch_docids", "ds2i_gpu_batch_prepare", "ds2i_gpu_batch_run", "ds2i_gpu_batch_run_ex", "ds2i_gpu_batch_fetch", "ds2i_gpu_batch_fetch_docids",
    "ds2i_gpu_batch_stats", "ds2i_gpu_batch_device_results", "ds2i_gpu_batch_device_docids", "ds2i_gpu_merge_shards", "ds2i_gpu_batch_free",
    "ds2i_gpu_decode_lists", "ds2i_gpu_next_geq_batch",
    "ds2i_gpu_decode_lists_checksum", "ds2i_gpu_index_list_bytes", "ds2i_gpu_index_type_known", "ds2i_gpu_batch_wait", "ds2i_gpu_batch_device_fused",
    "ds2i_gpu_group_open", "ds2i_gpu_group_close", "ds2i_gpu_group_size", "ds2i_gpu_group_index", "ds2i_gpu_group_query_batch",
]

_lib = None


class Ds2iGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ds2i_gpu error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not built: run `python -m ds2i_b200.build` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32p, u64p, f32p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
    L.ds2i_gpu_last_error.restype = C.c_char_p
    L.ds2i_gpu_op_from_name.argtypes = [C.c_char_p]
    L.ds2i_gpu_index_open.argtypes = [vp, C.c_size_t, C.c_char_p, C.c_int, C.POINTER(vp)]
    L.ds2i_gpu_index_open_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
    L.ds2i_gpu_index_close.argtypes = [vp]
    L.ds2i_gpu_index_close.restype = None
    for f in (L.ds2i_gpu_index_size, L.ds2i_gpu_index_num_docs, L.ds2i_gpu_index_device_bytes):
        f.argtypes = [vp]
        f.restype = C.c_uint64
    L.ds2i_gpu_index_set_global_stats.argtypes = [vp, u64p, C.c_size_t, C.c_uint64]
    L.ds2i_gpu_index_list_sizes.argtypes = [vp, u32p, C.c_size_t, u64p]
    L.ds2i_gpu_index_list_bytes.argtypes = [vp, u32p, C.c_size_t, u64p]
    L.ds2i_gpu_wand_open.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.ds2i_gpu_wand_open_file.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.ds2i_gpu_wand_close.argtypes = [vp]
    L.ds2i_gpu_wand_close.restype = None
    L.ds2i_gpu_query_batch.argtypes = [vp, vp, C.c_int, C.c_uint32, u32p, u64p, C.c_size_t, u64p, f32p, f32p]
    L.ds2i_gpu_query_batch_docids.argtypes = [vp, vp, C.c_int, C.c_uint32, u32p, u64p, C.c_size_t, u64p, f32p, u32p, f32p]
    L.ds2i_gpu_batch_prepare.argtypes = [vp, vp, u32p, u64p, C.c_size_t, C.POINTER(vp)]
    L.ds2i_gpu_batch_run.argtypes = [vp, C.c_int, C.c_uint32, f32p]
    L.ds2i_gpu_batch_run_ex.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, f32p]
    L.ds2i_gpu_batch_fetch.argtypes = [vp, u64p, f32p]
    L.ds2i_gpu_batch_fetch_docids.argtypes = [vp, u32p]
    L.ds2i_gpu_batch_stats.argtypes = [vp, u64p]
    L.ds2i_gpu_batch_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.ds2i_gpu_batch_device_docids.argtypes = [vp, C.POINTER(vp)]
    L.ds2i_gpu_merge_shards.argtypes = [vp, vp, vp, C.c_uint32, C.c_size_t, C.c_uint32, C.c_int, vp, vp, vp]
    L.ds2i_gpu_batch_free.argtypes = [vp]
    L.ds2i_gpu_batch_free.restype = None
    L.ds2i_gpu_decode_lists.argtypes = [vp, u32p, C.c_size_t, u64p, u32p, u32p, f32p]
    L.ds2i_gpu_next_geq_batch.argtypes = [vp, u32p, C.c_size_t, u64p, u64p, u64p, u64p, f32p]
    L.ds2i_gpu_decode_lists_checksum.argtypes = [vp, u32p, C.c_size_t, u64p, u64p, u64p, f32p]
    L.ds2i_gpu_index_type_known.argtypes = [C.c_char_p]
    L.ds2i_gpu_batch_wait.argtypes = [vp, f32p]
    L.ds2i_gpu_batch_device_fused.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.ds2i_gpu_group_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.ds2i_gpu_group_close.argtypes = [vp]
    L.ds2i_gpu_group_close.restype = None
    L.ds2i_gpu_group_size.argtypes = [vp]
    L.ds2i_gpu_group_index.argtypes = [vp, C.c_int]
    L.ds2i_gpu_group_index.restype = vp
    L.ds2i_gpu_group_query_batch.argtypes = [vp, C.c_int, C.c_uint32, u32p, u64p, C.c_size_t, u64p, f32p, u32p, f32p]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise Ds2iGpuError(rc, lib().ds2i_gpu_last_error().decode("utf-8", "replace"))
