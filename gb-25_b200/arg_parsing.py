"""CLI arguments of the GB-25 run scripts — host-side mirror of /root/reference/src/arg_parsing.jl.

``parse_baroclinic_instability_args`` (arg_parsing.jl:9-47) returns a dict with the reference's keys ("grid-x", "grid-y",
"grid-z", "float-type", "target-float-type", "limbs", "dimension"); ``float_type_from_args`` (:75-77) maps the string to a
type.  libgb25cuda is built for Float32 (libgb25cuda.so, all kernel generations) and Float64 (libgb25cuda_f64.so, the
operator-per-kernel generation): ``supported_float_type`` is what the run scripts call before building a model; the
multifloat lowering (:96-104) is a Reactant feature and has no counterpart."""
from __future__ import annotations

import argparse

import numpy as np

_FLOAT_TYPES = {"Float64": np.float64, "f64": np.float64, "Float32": np.float32, "f32": np.float32,
                "Float16": np.float16, "f16": np.float16, "BFloat16": "bfloat16", "bf16": "bfloat16",
                "Float8E5M2": "float8_e5m2", "f8E5M2": "float8_e5m2", "f8": "float8_e5m2",
                "Float8E4M3": "float8_e4m3", "f8E4M3": "float8_e4m3"}
_SHORT = {np.float64: "f64", np.float32: "f32", np.float16: "f16", "bfloat16": "bf16", "float8_e5m2": "f8E5M2",
          "float8_e4m3": "f8E4M3"}


def parse_baroclinic_instability_args(*, grid_x_default: int, grid_y_default: int, grid_z_default: int, argv=None) -> dict:
    """Same options, defaults and key names as the reference (the default float type is Float64 there too)."""
    p = argparse.ArgumentParser()
    p.add_argument("--grid-x", type=int, default=grid_x_default, help="Base factor for number of grid points on the x axis.")
    p.add_argument("--grid-y", type=int, default=grid_y_default, help="Base factor for number of grid points on the y axis.")
    p.add_argument("--grid-z", type=int, default=grid_z_default, help="Base factor for number of grid points on the z axis.")
    p.add_argument("--float-type", type=str, default="Float64",
                   help="The default Oceananigans float type (Float64/f64, Float32/f32, Float16/f16, BFloat16/bf16)")
    p.add_argument("--target-float-type", type=str, default="",
                   help="the float type for execution, or the empty string for no lowering")
    p.add_argument("--limbs", type=int, default=2, help="Number of lower-precision limbs in the multifloat lowering")
    p.add_argument("--dimension", type=str, default="first", help="Multifloat expansion dimension (first, last, tuple)")
    ns = p.parse_args(argv)
    return {"grid-x": ns.grid_x, "grid-y": ns.grid_y, "grid-z": ns.grid_z, "float-type": ns.float_type,
            "target-float-type": ns.target_float_type, "limbs": ns.limbs, "dimension": ns.dimension}


def float_type_from_string(s: str):
    try:
        return _FLOAT_TYPES[s]
    except KeyError:
        raise AssertionError(f"Unknown float type {s}") from None      # the reference throws an AssertionError


def float_type_from_args(parsed_args: dict):
    return float_type_from_string(parsed_args["float-type"])


def float_type_to_string(t) -> str:
    try:
        return _SHORT[t]
    except (KeyError, TypeError):
        raise AssertionError(f"Unknown float type {t}") from None


def multifloat_from_args(parsed_args: dict):
    """``nothing`` without a target type, as in the reference; the lowering itself is Reactant's and is not offered."""
    if parsed_args["target-float-type"] == "":
        return None
    float_type_from_string(parsed_args["float-type"]); float_type_from_string(parsed_args["target-float-type"])
    raise NotImplementedError("multifloat lowering (Reactant.MultiFloatOptions) has no counterpart in libgb25cuda")


def supported_float_type(parsed_args: dict):
    """Float32 or Float64 (the two builds of the library); anything else is refused loudly instead of being silently cast."""
    t = float_type_from_args(parsed_args)
    if t is not np.float32 and t is not np.float64:
        raise ValueError(f"libgb25cuda is built for Float32 and Float64; --float-type {parsed_args['float-type']} is not "
                         "available")
    return t
