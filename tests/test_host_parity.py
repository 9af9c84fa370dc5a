"""Host logic on the CPU: the product's Python mirror (constructors, kinematics, flattener) against the
oracle's independent C++ construction -- poses must agree bit-for-bit, since the finite-difference
normals of the reference amplify pose noise by 1e8."""
import math

import numpy as np
import pytest

from tests import scenes


def _prims(flat):
    return [flat._prims[i] for i in range(flat.n_prims)]


def test_doublet_prim_poses_bit_identical(bmo, orc):
    from bmo_b200.flatten import FlatSystem
    for rotate in (False, True):
        sc, osc = scenes.doublet_spot(bmo, rotate), scenes.doublet_spot_oracle(rotate)
        flat = FlatSystem(sc["system"], [707e-9])
        assert (flat.n_prims, flat.n_parts, flat.n_objects, flat.n_meshes) == (5, 3, 2, 1)
        # oracle poses of the two lens unions; members share the union pose only before rotation,
        # so compare the union handle poses with the host UnionSDF poses
        for k, lens in enumerate((sc["doublet"].front, sc["doublet"].back)):
            po, do = osc["doublet"].part(k).shape().pose()
            assert np.array_equal(np.array(lens.shape.pos), po)
            assert np.array_equal(np.array(lens.shape.dir), do)
        po, do = osc["spot"].pose()
        assert np.array_equal(np.array(sc["spot"].position()), po) and np.array_equal(np.array(sc["spot"].orientation()), do)
        ov = osc["spot"].shape().eval("mesh_vertices", nout=12)
        assert np.array_equal(sc["spot"].shape.vertices.ravel(), ov)


def test_michelson_poses_bit_identical(bmo, orc):
    sc, osc = scenes.michelson(bmo, pd_n=8), scenes.michelson_oracle(pd_n=8)
    for k in ("m1", "m2", "rpm", "pd"):
        po, do = osc[k].pose()
        assert np.array_equal(np.array(sc[k].position()), po), k
        assert np.array_equal(np.array(sc[k].orientation()), do), k
    for i, part in enumerate(sc["cbs"].parts):
        po, do = osc["cbs"].part(i).pose()
        assert np.array_equal(np.array(part.position()), po) and np.array_equal(np.array(part.orientation()), do)
    ov = osc["cbs"].part(2).shape().eval("mesh_vertices", nout=12)
    assert np.array_equal(sc["cbs"].coating.shape.vertices.ravel(), ov)


def test_flattener_order_and_roles(bmo):
    from bmo_b200 import flatten as fl
    sc = scenes.michelson(bmo, pd_n=8)
    flat = fl.FlatSystem(sc["system"], [632.8e-9])
    kinds = [flat._objs[i].kind for i in range(flat.n_objects)]
    assert kinds == [fl.OBJ_MIRROR, fl.OBJ_CUBE_BS, fl.OBJ_MIRROR, fl.OBJ_MIRROR, fl.OBJ_PD]   # Leaves() pre-order
    cube = flat._objs[1]
    roles = [flat._parts[cube.first_part + k].role for k in range(cube.n_parts)]
    assert roles == [fl.ROLE_FRONT, fl.ROLE_BACK, fl.ROLE_COATING]
    assert flat._ntab.shape == (2, 1) and flat._ntab[0, 0] == scenes.N_NBK7_6328
    # bounding spheres enclose the shapes (sampled surface points of the mirrors)
    for i in range(flat.n_parts):
        assert flat._parts[i].bound[3] > 0


def test_unknown_wavelength_raises_keyerror(bmo):
    from bmo_b200.flatten import FlatSystem
    n = bmo.DiscreteRefractiveIndex([632.8e-9], [1.5])
    lens = bmo.SphericalLens(0.1, -0.1, 5e-3, bmo.inch, n)
    with pytest.raises(KeyError):      # RefractiveIndexUtils.jl:31
        FlatSystem(bmo.System([lens]), [500e-9])


def test_nonintersectable_objects_are_skipped(bmo):
    from bmo_b200.flatten import FlatSystem
    dummy = bmo.NonInteractableObject(bmo.CubeMesh(0.1))
    m = bmo.RoundPlanoMirror(bmo.inch, 5e-3)
    flat = FlatSystem(bmo.System([dummy, bmo.ObjectGroup([m])]), [1e-6])
    assert flat.n_objects == 1 and flat.objects[0] is m


def test_sources_match_reference_formulas(bmo):
    src = bmo.UniformDiscSource((0.0, -0.05, 0.0), (0.0, 1.0, 0.0), 20e-3, 707e-9, num_rays=1000, e1=(1.0, 0.0, 0.0))
    pos, d = scenes.fibonacci_disc(1000)
    assert np.allclose(src.pos, pos, atol=1e-18) and np.array_equal(src.dir, d)
    cs = bmo.CollimatedSource((0, 0, 0), (0, 1, 0), 10e-3, 1e-6, num_rings=5, num_rays=200, b1=(1, 0, 0))
    assert len(cs) == 200 and np.allclose(np.linalg.norm(cs.pos[-1]), 5e-3)
    ps = bmo.PointSource((0, 0, 0), (0, 1, 0), 0.1, 1e-6, num_rings=5, num_rays=200, b1=(1, 0, 0))
    assert len(ps) == 200 and np.allclose(np.linalg.norm(ps.dir, axis=1), 1.0)
