"""Host-side mirror of /root/reference/src/arg_parsing.jl: same keys, defaults and error behaviour."""
import numpy as np
import pytest

from gb25_b200 import arg_parsing as A, sharding


def test_defaults_and_keys_follow_the_reference():
    d = A.parse_baroclinic_instability_args(grid_x_default=1536, grid_y_default=768, grid_z_default=4, argv=[])
    assert d == {"grid-x": 1536, "grid-y": 768, "grid-z": 4, "float-type": "Float64", "target-float-type": "",
                 "limbs": 2, "dimension": "first"}
    d = A.parse_baroclinic_instability_args(grid_x_default=1, grid_y_default=1, grid_z_default=1,
                                            argv=["--grid-x", "192", "--grid-y", "96", "--grid-z", "50", "--float-type", "f32"])
    assert (d["grid-x"], d["grid-y"], d["grid-z"]) == (192, 96, 50)
    assert A.float_type_from_args(d) is np.float32 and A.supported_float_type(d) is np.float32
    assert A.multifloat_from_args(d) is None


def test_float_type_strings():
    assert A.float_type_from_string("Float64") is np.float64 and A.float_type_from_string("f16") is np.float16
    assert A.float_type_from_string("bf16") == "bfloat16" and A.float_type_to_string(np.float32) == "f32"
    assert A.float_type_to_string(A.float_type_from_string("f8")) == "f8E5M2"
    with pytest.raises(AssertionError):
        A.float_type_from_string("Float128")
    with pytest.raises(AssertionError):
        A.float_type_to_string(int)


def test_float32_and_float64_are_built():
    d = A.parse_baroclinic_instability_args(grid_x_default=8, grid_y_default=8, grid_z_default=2, argv=[])
    assert A.supported_float_type(d) is np.float64      # the reference's default
    d["float-type"] = "f16"
    with pytest.raises(ValueError, match="Float32 and Float64"):
        A.supported_float_type(d)
    d["target-float-type"] = "f32"
    with pytest.raises(NotImplementedError):
        A.multifloat_from_args(d)


def test_tile_arithmetic_of_the_sharded_script():
    # sharding/sharded_baroclinic_instability_simulation_run.jl:82-88 with the CLI defaults of the scaling test
    Rx, Ry = sharding.factors(8)
    Nx, Ny = sharding.global_size_from_tile(1536, 768, Rx, Ry)
    assert (Nx, Ny) == (6128, 1520) and Nx % Rx == 0 and Ny % Ry == 0
