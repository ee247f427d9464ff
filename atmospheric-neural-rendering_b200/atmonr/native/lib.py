"""ctypes binding of libatmonr_b200.so (the C ABI declared in include/atmonr_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised. Tensors are passed as raw device pointers together with the current CUDA stream.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG_ROOT = Path(__file__).resolve().parents[2]
LIB_PATH = Path(os.environ.get("ATMONR_B200_LIB", _PKG_ROOT / "lib" / "libatmonr_b200.so"))

MAX_LEVELS = 16


class GridT(C.Structure):
    _fields_ = [
        ("n_dims", C.c_int32), ("n_levels", C.c_int32), ("n_feat", C.c_int32), ("reserved", C.c_int32),
        ("scale", C.c_float * MAX_LEVELS), ("res", C.c_uint32 * MAX_LEVELS),
        ("size", C.c_uint32 * MAX_LEVELS), ("offset", C.c_uint32 * (MAX_LEVELS + 1)),
    ]

    @property
    def n_entries(self) -> int:
        return int(self.offset[self.n_levels])


class FrameT(C.Structure):
    _fields_ = [
        ("scale", C.c_double), ("offset", C.c_double * 3),
        ("lat_min", C.c_double), ("lat_range", C.c_double), ("lon_min", C.c_double), ("lon_range", C.c_double),
        ("origin_height", C.c_double), ("shift_lon", C.c_int32), ("enabled", C.c_int32),
    ]


class MlpT(C.Structure):
    _fields_ = [
        ("n_in", C.c_int32), ("in_pad", C.c_int32), ("width", C.c_int32),
        ("n_hidden", C.c_int32), ("n_out", C.c_int32), ("out_pad", C.c_int32),
    ]

    @property
    def n_params(self) -> int:
        return self.width * self.in_pad + (self.n_hidden - 1) * self.width * self.width + self.out_pad * self.width


P, I64, I32, U64, F32, F64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_float, C.c_double
GP, FP, MP = C.POINTER(GridT), C.POINTER(FrameT), C.POINTER(MlpT)

# name -> argtypes; must list EVERY function include/atmonr_b200.h declares (checked by tests)
SIGNATURES = {
    "atmonr_abi_version": [],
    "atmonr_last_error": [],
    "atmonr_l2_persist": [P, C.c_size_t, F32, P],
    "atmonr_grid_layout": [I32, I32, I32, I32, F32, GP],
    "atmonr_get_rays": [P, P, P, P, P, I64, F32, F64, I32, P, P, P, P, C.POINTER(C.c_int), P],
    "atmonr_filter_rays": [P, P, P, I64, P, P],
    "atmonr_ray_extent": [P, P, P, I64, P, P, P],
    "atmonr_normalize_origins": [P, I64, P, F64, P, P],
    "atmonr_gather_batch": [P, P, P, P, P, P, P, P, I64, I64, P, P, P, P, P, P, P, P, P],
    "atmonr_sample_uniform": [P, P, P, P, P, I64, I32, I32, U64, U64, P, P, P],
    "atmonr_preprocess_horizontal": [FP, P, P, I64, I32, P],
    "atmonr_ngp_sample_points": [FP, P, P, P, P, P, I64, I32, I32, U64, U64, F32, P, P, P],
    "atmonr_ngp_sample_points_height": [FP, P, P, P, P, P, I64, I32, I32, U64, U64, F32, F64, C.POINTER(C.c_double), F64, P, P, P],
    "atmonr_hashgrid_fwd": [GP, P, I32, P, I64, P, P],
    "atmonr_hashgrid_bwd": [GP, P, I32, P, I64, P, P],
    "atmonr_hashgrid_indices": [GP, P, I32, I64, P, P],
    "atmonr_mlp_fwd": [MP, P, P, I64, P, P],
    "atmonr_mlp_bwd": [MP, P, P, P, I64, P, P, P],
    "atmonr_ngp_field_fwd": [GP, P, MP, P, MP, P, P, P, I64, I32, P, P, P],
    "atmonr_ngp_field_bwd": [GP, P, MP, P, MP, P, P, P, P, P, I64, I32, P, P, P, P],
    "atmonr_ngp_surface_fwd": [GP, P, MP, P, P, P, P, I64, P, P],
    "atmonr_ngp_surface_bwd": [GP, P, MP, P, P, P, P, P, I64, P, P, P],
    "atmonr_composite_fwd": [P, P, P, P, F32, I64, I32, I32, I32, I32, P, P, P, P, P, P, P],
    "atmonr_composite_bwd": [P, P, P, P, P, P, P, P, F32, I64, I32, I32, I32, I32, P, P, P, P, P, P],
    "atmonr_composite_bwd_compact": [P, P, P, P, P, P, P, P, F32, I64, I32, I32, I32, I32, P, P, P, P, P, P, P],
    "atmonr_composite_bwd_weights": [P, P, P, P, P, P, P, P, P, P, F32, I64, I32, I32, I32, I32, P, P, P, P, P],
    "atmonr_band_loss": [P, P, P, F32, I32, I64, I32, F32, P, P, P, P],
    "atmonr_adamw_step": [P, P, P, P, P, I64, F64, F64, F64, F64, F64, I64, F64, I32, P],
    "atmonr_extract_sigma": [FP, GP, P, MP, P, P, I64, F32, P, P],
    "atmonr_extract_sigma_tc": [FP, GP, P, MP, P, P, I64, F32, P, P],
    "atmonr_positional_encoding": [P, I64, I32, C.POINTER(C.c_int32), I32, P, P],
    "atmonr_positional_encoding_f64": [P, I64, I32, C.POINTER(C.c_int32), I32, P, P],
    "atmonr_sample_pdf": [P, P, P, I64, I32, I32, P, P, P],
    "atmonr_sample_pdf_train": [P, P, P, I64, I32, I32, P, P, P, P, P],
    "atmonr_sample_pdf_bwd": [P, P, P, P, P, P, P, I64, I32, I32, P, P, P],
    "atmonr_nerf_encode": [FP, P, P, P, I64, I32, C.POINTER(C.c_int32), I32, P, I32, P, P],
    "atmonr_nerf_encode_bwd": [FP, P, P, P, P, P, I32, I64, I32, C.POINTER(C.c_int32), P, P],
    "atmonr_composite_dz": [P, I64, I32, F32, P, P],
    "atmonr_append_heights": [P, I64, F64, C.POINTER(C.c_double), F64, P, P],
    "atmonr_tc_probe": [P, P, I32, P, P],
    "atmonr_linear_prep": [P, I32, I32, I32, I32, P, P],
    "atmonr_linear_fwd_tc": [P, I64, P, I64, I32, P, I64, P, P, I64, I32, I32, I32, I32, P, I64, P, P, P, I64, P],
    "atmonr_linear_dw_tc": [P, I64, P, I64, P, I64, P, I64, I32, I64, I32, I32, I32, P, P, P],
    "atmonr_ngp_field_fwd_tc": [GP, P, MP, P, MP, P, P, P, I64, I32, P, P, P, P],
    "atmonr_ngp_field_bwd_tc": [GP, P, MP, P, MP, P, P, P, P, P, P, P, I64, I32, P, P, P, P],
    "atmonr_ngp_field_bwd_tc_compact": [GP, MP, P, MP, P, P, P, P, P, P, P, P, P, I64, I32, P, P, P, P],
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python atmospheric-neural-rendering_b200/build.py` "
            "(or __graft_entry__.build()). There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "atmonr_last_error" else C.c_int
    _lib = lib
    return lib


# kernels launched by each entry point (for launch accounting in bench.py)
LAUNCHES = {"atmonr_band_loss": 2, "atmonr_grid_layout": 0, "atmonr_abi_version": 0, "atmonr_last_error": 0,
            "atmonr_l2_persist": 0, "atmonr_ray_extent": 2}


class CallStats:
    """Optional per-entry-point accounting: launch counts and, when `timed`, CUDA-event durations
    recorded on the launching stream (bench.py reads them after a synchronize)."""

    def __init__(self, timed: bool = False):
        self.timed = timed
        self.launches = 0
        self.calls: dict[str, int] = {}
        self.events: dict[str, list] = {}

    def durations_ms(self) -> dict[str, list[float]]:
        torch.cuda.synchronize()
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in self.events.items()}


STATS: CallStats | None = None


def call(name: str, *args) -> None:
    lib = load()
    st = STATS
    if st is not None:
        st.launches += LAUNCHES.get(name, 1)
        st.calls[name] = st.calls.get(name, 0) + 1
        if st.timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = getattr(lib, name)(*args)
            b.record()
            st.events.setdefault(name, []).append((a, b))
        else:
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise NativeLibraryError(f"{name} failed ({rc}): {lib.atmonr_last_error().decode()}")


def ptr(t: torch.Tensor | None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeLibraryError("libatmonr_b200 operates on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise NativeLibraryError("non-contiguous tensor passed to libatmonr_b200")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def grid_layout(n_dims: int, cfg: dict) -> GridT:
    """Level table of a HashGrid encoding config (host-only call; works without a GPU)."""
    if int(cfg.get("n_features_per_level", 2)) != 2:
        raise NativeLibraryError("only n_features_per_level == 2 is supported")
    g = GridT()
    call(
        "atmonr_grid_layout", n_dims, int(cfg["n_levels"]), int(cfg["log2_hashmap_size"]),
        int(cfg["base_resolution"]), float(cfg["per_level_scale"]), C.byref(g),
    )
    return g


def mlp_shape(n_in: int, n_out: int, cfg: dict) -> MlpT:
    if cfg.get("otype", "FullyFusedMLP") not in ("FullyFusedMLP", "CutlassMLP"):
        raise NativeLibraryError(f"unsupported network otype {cfg.get('otype')}")
    if cfg.get("activation", "ReLU") != "ReLU" or cfg.get("output_activation", "None") != "None":
        raise NativeLibraryError("only ReLU hidden activation and no output activation are supported")
    m = MlpT()
    m.n_in, m.in_pad = n_in, (n_in + 15) // 16 * 16
    m.width, m.n_hidden = int(cfg["n_neurons"]), int(cfg["n_hidden_layers"])
    m.n_out, m.out_pad = n_out, (n_out + 15) // 16 * 16
    return m


def make_frame(scale, offset, lat_min, lat_range, lon_min, lon_range, origin_height, shift_lon) -> FrameT:
    f = FrameT()
    f.scale = float(scale)
    for k in range(3):
        f.offset[k] = float(offset[k])
    f.lat_min, f.lat_range = float(lat_min), float(lat_range)
    f.lon_min, f.lon_range = float(lon_min), float(lon_range)
    f.origin_height = float(origin_height)
    f.shift_lon, f.enabled = int(bool(shift_lon)), 1
    return f


def disabled_frame() -> FrameT:
    return FrameT()
