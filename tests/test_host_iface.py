"""Host-side mirror of fft.generate_fft_interface: type selection, region conventions and the
assertions the reference makes (no GPU needed: nothing is executed)."""
import numpy as np
import pytest
import torch


def test_type_selection_follows_reference(fft):
    # src/fft.rg:231-243: (sizeof(dtype_out), real_flag) -> cufftType
    lib = fft._lib
    assert fft.generate_fft_interface(fft.int1d, fft.complex64, fft.complex64).ftype == lib.Z2Z
    assert fft.generate_fft_interface(fft.int3d, fft.double, fft.complex64).ftype == lib.D2Z
    assert fft.generate_fft_interface(fft.int2d, fft.complex32, fft.complex32).ftype == lib.C2C
    assert fft.generate_fft_interface(fft.int1d, fft.float32, fft.complex32).ftype == lib.R2C
    assert fft.generate_fft_interface(fft.int1d, "double", "complex64").real_flag


def test_generate_rejects_bad_arguments(fft):
    with pytest.raises(AssertionError):
        fft.generate_fft_interface(3, fft.complex64, fft.complex64)   # not an index type (src/fft.rg:32)
    with pytest.raises(TypeError):
        fft.generate_fft_interface(fft.int1d, "quad", fft.complex64)


def test_interface_table_has_the_reference_entries(fft):
    iface = fft.generate_fft_interface(fft.int1d, fft.complex64, fft.complex64)
    for name in ("plan", "make_plan", "make_plan_batch", "make_plan_task", "make_plan_distrib", "execute_plan",
                 "execute_plan_task", "destroy_plan", "destroy_plan_task", "destroy_plan_distrib", "get_plan",
                 "get_tunable", "get_num_nodes", "get_num_local_gpus"):
        assert hasattr(iface, name), name


def test_region_layout_is_legion_default(fft):
    r = fft.Region((3, 3, 2), fft.complex64, device="cpu")
    assert r.bounds == ((0, 0, 0), (2, 2, 1)) and r.extent == (3, 3, 2) and r.volume == 18
    assert r.offsets == (16, 48, 144)
    assert r.offsets[2] // r.offsets[0] == 9              # i_dist of make_plan_batch (src/fft.rg:372-377)
    d = fft.Region((4, 2), fft.double, device="cpu")
    assert d.offsets == (8, 32)


def test_bounds_and_type_assertions(fft):
    iface = fft.generate_fft_interface(fft.int1d, fft.complex64, fft.complex64)
    r = fft.Region((8,), fft.complex64, device="cpu")
    s = fft.Region((4,), fft.complex64, device="cpu")
    p = fft.PlanRegion(1)
    with pytest.raises(AssertionError, match="identical in size"):   # src/fft.rg:276
        iface.make_plan(r, s, p)
    s2 = fft.Region((8,), fft.complex32, device="cpu")
    with pytest.raises(AssertionError, match="element types"):
        iface.make_plan(r, s2, p)
    s3 = fft.Region((8,), fft.complex64, device="cpu")
    with pytest.raises(AssertionError, match="no CPU path"):         # fail loudly: no fallback
        iface.make_plan(r, s3, p)


def test_plan_region_is_pod_and_node_checked(fft):
    iface = fft.generate_fft_interface(fft.int1d, fft.complex64, fft.complex64)
    p = fft.PlanRegion(1)
    assert p.data.dtype == iface.plan and p.data.itemsize == 32
    q = iface.get_plan(p, False)
    q["b200_p"][0] = 77
    q["address_space"][0] = 5                                       # made on another node
    with pytest.raises(AssertionError, match="node where"):         # src/fft.rg:186
        iface.get_plan(p, True)
    copy = np.copy(p.data)                                          # Legion may memcpy the plan region
    assert int(copy["b200_p"][0]) == 77


def test_partition_equal(fft):
    r = fft.Region((6,), fft.complex64, device="cpu")
    r.flat[:] = torch.arange(6, dtype=torch.float64).to(torch.complex128)
    parts = r.partition_equal(2)
    assert [q.volume for q in parts] == [3, 3] and parts[1].bounds == ((3,), (5,))
    parts[1].flat[0] = -1
    assert r.flat[3] == -1                                          # views, not copies
    pp = fft.PlanRegion(2).partition_equal(2)
    assert len(pp) == 2 and pp[0].volume == 1


def test_tunables_without_distributed(fft):
    iface = fft.generate_fft_interface(fft.int1d, fft.complex64, fft.complex64)
    assert iface.get_num_nodes() == 1
    assert iface.get_num_local_gpus() == (torch.cuda.device_count() if torch.cuda.is_available() else 0)
    assert iface.packed_output_shape((4, 6)) == (4, 6)
    riface = fft.generate_fft_interface(fft.int2d, fft.double, fft.complex64)
    assert riface.packed_output_shape((4, 6)) == (4, 4)
