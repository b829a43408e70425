"""CPU: the host-side mirror of the reference interface (data model, ordering, errors)."""
import numpy as np
import pytest


def test_point_order_boundary_then_volume(pkg):
    b1, b2, v = np.random.rand(4, 3), np.random.rand(3, 3), np.random.rand(5, 3)
    cloud = pkg.PointCloud(pkg.PointBoundary({"inlet": b1, "wall": b2}), v)
    p = pkg.points(cloud)                                   # src/cloud.jl:235-237, src/boundary.jl:164
    assert len(cloud) == 12 and np.array_equal(p, np.concatenate([b1, b2, v]))


def test_machine_type_promotion(pkg):
    cloud = pkg.PointCloud(np.random.rand(4, 3).astype(np.float32), np.random.rand(5, 3))   # src/cloud.jl:38-56
    assert pkg.points(cloud).dtype == np.float64
    cloud32 = pkg.PointCloud(np.random.rand(4, 2).astype(np.float32), np.random.rand(5, 2).astype(np.float32))
    assert pkg.points(cloud32).dtype == np.float32


def test_notopology_default_and_errors(pkg):
    cloud = pkg.PointCloud(np.random.rand(10, 3))
    assert isinstance(pkg.topology(cloud), pkg.NoTopology) and not pkg.hastopology(cloud)   # test/topology.jl:1-10
    with pytest.raises(pkg.WtpArgumentError):               # test/topology.jl:96-104 (ArgumentError)
        pkg.neighbors(cloud)
    with pytest.raises(pkg.WtpArgumentError):
        pkg.neighbors(cloud, 1)
    assert pkg.rebuild_topology_(cloud) is None             # no-op (test/topology.jl:86-94)
    assert repr(pkg.NoTopology()) == "NoTopology()"         # test/topology.jl:122-126


def test_topology_storage_and_printing(pkg):
    rows = pkg.FlatRows(np.arange(12, dtype=np.int64).reshape(4, 3) + 1)
    t = pkg.KNNTopology(rows, 3)
    assert len(pkg.neighbors(t)) == 4 and pkg.neighbors(t, 2).tolist() == [4, 5, 6]
    assert "KNNTopology" in t.show() and "k: 3" in t.show() and repr(t) == "KNNTopology(k=3)"   # test/topology.jl:117-118
    csr = pkg.CSRRows(np.array([0, 2, 2, 5]), np.array([2, 3, 1, 2, 4]))
    r = pkg.RadiusTopology(csr, 0.15)
    assert [x.tolist() for x in pkg.neighbors(r)] == [[2, 3], [], [1, 2, 4]] and r.radius == 0.15


def test_repel_argument_validation_precedes_device_work(pkg):
    cloud = pkg.PointCloud(np.random.rand(10, 3), np.random.rand(20, 3))
    sp = pkg.ConstantSpacing(0.1)
    with pytest.raises(pkg.WtpArgumentError):               # src/repel.jl:74
        pkg.repel(cloud, sp, rebuild_every=0)
    with pytest.raises(pkg.WtpError):                       # arbitrary callables cannot cross the ABI
        pkg.repel(cloud, lambda x: 0.1)
    class MyForce(pkg.RepelForceModel):
        kind = "mine"
    with pytest.raises(pkg.WtpError):
        pkg.repel(cloud, sp, force_model=MyForce())
    with pytest.raises(pkg.WtpError):
        pkg.repel(cloud, sp, octree=object())


def test_spacing_constructors_validate(pkg):
    with pytest.raises(pkg.WtpArgumentError):               # spacings.jl:61-62
        pkg.LogLike(np.zeros((0, 3)), 0.1, 1.5)
    with pytest.raises(pkg.WtpArgumentError):               # spacings.jl:103-104
        pkg.BoundaryLayerSpacing(np.random.rand(5, 3), at_wall=0.1, bulk=1.0, layer_thickness=0.0)
    f = pkg.ClippedSpacingForce(0.5)
    assert f.u0 == 1.0 and pkg.StrongSpacingForce(0.5).gamma == 3.0 and pkg.InverseDistanceForce().beta == 0.2   # test/repel.jl:142-170


def test_sorting_networks_are_the_generators_output():
    """The register sorting networks compiled into the tiled kernels (csrc/sortnet*.inc) are exactly what
    scripts/gen_sortnet.py emits (the generator checks each network with the 0-1 principle and random permutations
    before it prints it)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gen = os.path.join(root, "scripts", "gen_sortnet.py")
    csrc = os.path.join(root, "whatsthepoint.jl_b200", "csrc")
    for name, args in (("sortnet48.inc", ("48", "33")), ("sortnet32.inc", ("32", "32")), ("sortnet48full.inc", ("48", "48"))):
        out = subprocess.run([sys.executable, gen, *args], capture_output=True, text=True, check=True).stdout
        assert out.strip() == open(os.path.join(csrc, name)).read().strip(), name
