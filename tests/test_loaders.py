"""Mesh import with the reference's names (import part): Utilities::TetGenLoader / OBJLoader semantics
(Utils/TetGenLoader.cpp, Utils/OBJLoader.h) on small committed files, and -- where the reference tree is present -- on its own
data/models/armadillo_4k.{node,ele}, whose tet model is then built by the host mirror and by the reference from the same arrays."""
import os
import numpy as np
import pytest

import positionbaseddynamics_b200.pypbd as pbd

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meshes")
X = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], dtype=np.float32)
T = np.array([0, 1, 2, 3, 1, 2, 3, 4], dtype=np.uint32)


def test_tet_formats():
    for x, t in (pbd.TetGenLoader.loadTetgenModel(os.path.join(HERE, "tiny.node"), os.path.join(HERE, "tiny.ele")),
                 pbd.TetGenLoader.loadTetFile(os.path.join(HERE, "tiny.tet")),
                 pbd.TetGenLoader.loadMSHModel(os.path.join(HERE, "tiny.msh"))):   # 1-based in the file, 0-based in memory
        assert x.dtype == np.float32 and t.dtype == np.uint32
        assert (x == X).all() and (t == T).all()


def test_obj_loader():
    x, normals, tex, faces = pbd.OBJLoader.loadObj(os.path.join(HERE, "tiny.obj"), (2.0, 1.0, 3.0))
    assert (x == np.array([[0, 0, 0], [2, 0, 0], [2, 1, 0], [0, 1, 0]], dtype=np.float32)).all()   # per-axis scale on the positions
    assert normals.shape == (1, 3) and tex.shape == (4, 2) and len(faces) == 2
    assert faces[1].posIndices == [0, 2, 3] and faces[1].texIndices == [0, 2, 3] and faces[1].normalIndices == [0, 0, 0]
    vd, mesh = pbd.OBJLoader.loadObjToMesh(os.path.join(HERE, "tiny.obj"), (1.0, 1.0, 1.0))
    assert vd.size() == 4 and mesh.numFaces() == 2 and (mesh.getFaces() == [0, 1, 2, 0, 2, 3]).all()
    # the loaded mesh goes straight into the model (pyPBD/examples/bunny_cloth.py style)
    pbd.Simulation._current = None
    sim = pbd.Simulation.getCurrent(); sim.initDefault(); model = sim.getModel()
    tm = model.addTriangleModel(vd.getVertices(), mesh.getFaces())
    assert tm.getParticleMesh().numFaces() == 2 and tm.getParticleMesh().numEdges() == 5


def test_armadillo_tet_model_like_the_reference(cpu_libs):
    node, ele = "/root/reference/data/models/armadillo_4k.node", "/root/reference/data/models/armadillo_4k.ele"
    if not (os.path.exists(node) and os.path.exists(ele)):
        pytest.skip("reference data files not present on this box")
    x, t = pbd.TetGenLoader.loadTetgenModel(node, ele)
    assert x.shape == (1180, 3) and len(t) == 4 * 3717
    from positionbaseddynamics_b200.model import HostModel
    from conftest import have_ref
    hm = HostModel(); other = cpu_libs.CpuPbd("ref" if have_ref("f64") else "oracle", "f64")
    for m in (hm, other):
        m.add_tet_model(x, t.reshape(-1, 4))
        m.add_solid_constraints(0, 2, k=1.0e6, nu=0.3)
    hm.init_groups(); other.init_groups()
    assert hm.num_constraints() == other.num_constraints() == 3717
    off_a, ids_a = hm.groups(); off_b, ids_b = other.groups()
    assert (off_a == off_b).all() and (ids_a == ids_b).all()          # same colouring of the imported mesh
    assert (hm.tet_edges(0) == other.tet_edges(0)).all()
    hm.close()
