"""ctypes binding of ``libmapdamage_b200.so`` (``include/mapdamage_b200.h``).

There is no Python or CPU implementation behind this module: if the CUDA
library is missing or cannot be loaded, importing the engine fails loudly.
"""
import ctypes as C
import os
from pathlib import Path

LIBRARY = Path(__file__).resolve().parent / "libmapdamage_b200.so"

ABI_VERSION = 1
N_CLASSES = 30

OK = 0
ERR_ARGUMENT, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE, ERR_CAPACITY, ERR_DATA, ERR_NCCL = range(-1, -8, -1)


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("mapdamage_b200 native error %d: %s" % (code, message))
        self.code = code
        self.message = message


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("length", C.c_int32), ("around", C.c_int32), ("min_qual", C.c_int32),
        ("n_libraries", C.c_int32), ("lg_bins", C.c_int32), ("n_slots", C.c_int32), ("reserved", C.c_int32),
        ("max_reads", C.c_int64), ("max_cigar_ops", C.c_int64), ("max_bases", C.c_int64),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_reads", C.c_int64), ("n_cigar", C.c_int64), ("n_bases", C.c_int64),
        ("flag", C.c_void_p), ("tid", C.c_void_p), ("pos", C.c_void_p), ("lib", C.c_void_p),
        ("l_seq", C.c_void_p), ("base_off", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p),
        ("seq4", C.c_void_p), ("qual", C.c_void_p), ("tlen", C.c_void_p), ("mtid", C.c_void_p),
        ("mpos", C.c_void_p),
    ]


class SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_reads", C.c_int64), ("len_lo", C.c_int32), ("len_hi", C.c_int32),
        ("mix", C.c_int32 * 4), ("paired", C.c_int32), ("with_qual", C.c_int32),
        ("n_libraries", C.c_int32), ("reserved", C.c_int32),
        ("error_rate", C.c_float), ("read_n_rate", C.c_float), ("filtered_rate", C.c_float),
        ("damage0", C.c_float), ("damage_decay", C.c_float), ("reserved2", C.c_float),
    ]


# every symbol include/mapdamage_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mdg_abi_version": (C.c_int, []),
    "mdg_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config)]),
    "mdg_destroy": (None, [C.c_void_p]),
    "mdg_last_error": (C.c_char_p, [C.c_void_p]),
    "mdg_device_pci_bus_id": (C.c_int, [C.c_int32, C.c_char_p, C.c_int32]),
    "mdg_host_alloc": (C.c_void_p, [C.c_size_t]),
    "mdg_host_free": (None, [C.c_void_p]),
    "mdg_set_reference": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]),
    "mdg_genome_composition": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdg_synth_reference": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint64]),
    "mdg_reference_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "mdg_count_submit": (C.c_int, [C.c_void_p, C.POINTER(Batch)]),
    "mdg_batch_upload": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(C.c_void_p)]),
    "mdg_batch_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdg_count_resident": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdg_sync": (C.c_int, [C.c_void_p]),
    "mdg_reset_tables": (C.c_int, [C.c_void_p]),
    "mdg_fetch_tables": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mdg_fetch_lg_overflow": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64]),
    "mdg_set_rescale_model": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "mdg_rescale_submit": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p]),
    "mdg_fetch_rescale_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdg_fetch_rescale_hist": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mdg_synth_batch": (C.c_int, [C.c_void_p, C.POINTER(SynthParams), C.POINTER(C.c_void_p)]),
    "mdg_batch_sizes": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64)]),
    "mdg_batch_download": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Batch)]),
    "mdg_bam_open": (C.c_int, [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "mdg_bam_close": (None, [C.c_void_p]),
    "mdg_bam_error": (C.c_char_p, [C.c_void_p]),
    "mdg_bam_header_text": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
    "mdg_bam_n_references": (C.c_int32, [C.c_void_p]),
    "mdg_bam_reference": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_uint32)]),
    "mdg_bam_set_libraries": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_int32]),
    "mdg_bam_read_batch": (C.c_int64, [C.c_void_p, C.POINTER(Batch), C.c_int64, C.c_int64, C.c_int64, C.c_uint32,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64)]),
    "mdg_bam_lenient_libraries": (C.c_int, [C.c_void_p, C.c_int32]),
    "mdg_bam_library_failure": (C.c_int64, [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]),
    "mdg_bam_records_seen": (C.c_int64, [C.c_void_p]),
    "mdg_bam_data_start": (C.c_uint64, [C.c_void_p]),
    "mdg_bam_stream_open": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int32, C.c_int64, C.POINTER(C.c_void_p)]),
    "mdg_bam_stream_close": (None, [C.c_void_p]),
    "mdg_bam_stream_error": (C.c_char_p, [C.c_void_p]),
    "mdg_bam_stream_set_libraries": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_int32]),
    "mdg_bam_stream_next": (C.c_int64, [C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "mdg_bam_stream_has_mr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "mdg_bam_stream_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "mdg_bam_write_raw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "mdg_bam_encode_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mdg_bam_encode_flush": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "mdg_rescale_submit_sparse": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "mdg_rescale_collect": (C.c_int64, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "mdg_rescale_resident": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mdg_inflate_raw": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "mdg_inflater_create": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p)]),
    "mdg_inflater_free": (None, [C.c_void_p]),
    "mdg_inflater_error": (C.c_char_p, [C.c_void_p]),
    "mdg_inflate_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mdg_sample_fraction": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, C.c_void_p]),
    "mdg_sample_reservoir": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "mdg_bam_create": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_int32, C.POINTER(C.c_void_p)]),
    "mdg_bam_writer_error": (C.c_char_p, [C.c_void_p]),
    "mdg_bam_write_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "mdg_bam_write_soa": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_int64, C.c_char_p, C.POINTER(C.c_char_p),
                                    C.c_int32]),
    "mdg_bam_finish": (C.c_int, [C.c_void_p]),
    "mdg_bam_writer_free": (None, [C.c_void_p]),
    "mdg_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "mdg_nccl_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "mdg_allreduce_tables": (C.c_int, [C.c_void_p]),
    "mdg_event_record": (C.c_int, [C.c_void_p, C.c_int32]),
    "mdg_event_elapsed_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mdg_launch_count": (C.c_int64, [C.c_void_p]),
    "mdg_last_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
}

_lib = None


def load():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIBRARY.is_file():
            raise ImportError(
                "%s is missing: build it with `python -m mapdamage_b200.build` "
                "(mapdamage_b200 has no CPU implementation)" % LIBRARY)
        lib = C.CDLL(os.environ.get("MDG_LIBRARY") or str(LIBRARY))  # MDG_LIBRARY: an alternative build (A/B runs)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.mdg_abi_version() != ABI_VERSION:
            raise ImportError("libmapdamage_b200.so has ABI %d, expected %d" % (lib.mdg_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def last_error(ctx):
    text = load().mdg_last_error(ctx)
    return text.decode("utf-8", "replace") if text else ""


def check(code, ctx=None):
    if code < 0:
        raise NativeError(code, last_error(ctx))
    return code
