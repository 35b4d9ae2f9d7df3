"""The pyPBD-named facade: scene construction in the reference's own call style (pyPBD/examples/cloth_model.py:18-124)."""
import math
import numpy as np
import pytest


def _build():
    import positionbaseddynamics_b200.pypbd as pbd
    pbd.Simulation._current = None
    sim = pbd.Simulation.getCurrent(); sim.initDefault(); model = sim.getModel()
    a = math.pi / 2
    R = [[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]]
    model.addRegularTriangleModel(20, 20, [0, 1, 0], R, [10, 10])
    pd = model.getParticles()
    pd.setMass(0, 0.0); pd.setMass(19, 0.0)
    tm = model.getTriangleModels()[0]
    model.addClothConstraints(tm, 4, 1.0e5, 1.0, 1.0, 1.0, 0.3, 0.3, False, False)
    model.addBendingConstraints(tm, 3, 100.0)
    return pbd, sim, model


def test_scene_construction_matches_oracle(cpu_libs):
    import scenes
    pbd, sim, model = _build()
    o = cpu_libs.CpuPbd("oracle", "f64")
    scenes.cloth(o, 20, 20, 4, 3, dist_k=1e5, bend_k=100.0)
    groups = model.getConstraintGroups()
    off, ids = o.groups()
    assert len(groups) == len(off) - 1 and all((groups[g] == ids[off[g]:off[g + 1]]).all() for g in range(len(groups)))
    assert model.numConstraints() == o.num_constraints()
    assert model.getTriangleModels()[0].getParticleMesh().numFaces() == 2 * 19 * 19
    pd = model.getParticles()
    assert pd.size() == 400 and pd.getMass(0) == 0.0 and pd.getInvMass(1) == 1.0
    assert np.allclose(pd.getVertices(), o.get("x"), atol=1e-6)
    pd.setPosition(5, [1, 2, 3]); assert (pd.getPosition(5) == [1, 2, 3]).all()
    c = model.getConstraints()[0]
    assert c["type"] == "Distance_XPBD" and len(c["bodies"]) == 2


@pytest.mark.gpu
def test_step_through_facade(cpu_libs):
    import scenes
    from parity_util import rel_position_error
    pbd, sim, model = _build()
    ts = sim.getTimeStep()
    ts.setValueUInt(pbd.TimeStepController.NUM_SUB_STEPS, 1)
    ts.setValueUInt(pbd.TimeStepController.MAX_ITERATIONS, 5)
    pbd.TimeManager.getCurrent().setTimeStepSize(0.005)
    o = cpu_libs.CpuPbd("oracle", "f64")
    scenes.cloth(o, 20, 20, 4, 3, dist_k=1e5, bend_k=100.0, sub_steps=1, max_iter=5)
    for _ in range(5):
        ts.step(model)
    o.step(5)
    assert rel_position_error(model.getParticles().getVertices(), o.get("x")) <= 1e-4
    assert abs(pbd.TimeManager.getCurrent().getTime() - 0.025) < 1e-6
    sim.reset()
    assert np.allclose(model.getParticles().getVertices(), o.get("x0"), atol=1e-6)


def test_builders_return_the_model_like_pypbd_and_refuse_collision_meshes():
    """pyPBD's addRegular*Model / add*Model return the new model (SimulationModelModule.cpp:98-229) and take testMesh; the
    reference's example scripts use both."""
    import positionbaseddynamics_b200.pypbd as pbd
    pbd.Simulation._current = None
    sim = pbd.Simulation.getCurrent(); sim.initDefault(); model = sim.getModel()
    tri = model.addRegularTriangleModel(4, 4, [0, 0, 0], np.eye(3), [1, 1], testMesh=False)
    tet = model.addRegularTetModel(3, 3, 3, testMesh=False)
    assert tri.getIndexOffset() == 0 and tri.getParticleMesh().numFaces() == 18
    assert tet.getIndexOffset() == 16 and tet.getParticleMesh().numTets() == 40
    tri2 = model.addTriangleModel([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [0, 1, 2])
    assert tri2.getIndexOffset() == 16 + 27 and tri2.getParticleMesh().numFaces() == 1
    with pytest.raises(pbd.PbdError):
        model.addRegularTriangleModel(4, 4, testMesh=True)
    pbd.Logger.addConsoleSink(pbd.LogLevel.INFO)
    pbd.Timing.reset(); pbd.Timing.printAverageTimes()


@pytest.mark.gpu
def test_example_scripts_run():
    """examples/cloth_model.py and beam_model.py (headless counterparts of the reference's pyPBD examples): pinned particles stay,
    everything stays finite, the cloth falls."""
    import importlib.util, os
    import positionbaseddynamics_b200.pypbd as pbd
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")
    out = {}
    for name in ("cloth_model", "beam_model"):
        pbd.Simulation._current = None
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, name + ".py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        out[name] = mod.main(frames=2)
        assert np.isfinite(out[name]).all()
        pd = pbd.Simulation.getCurrent().getModel().getParticles()
        pinned = [i for i in range(pd.size()) if pd.getMass(i) == 0.0]
        assert len(pinned) == (2 if name == "cloth_model" else 25)
        for i in pinned:                                   # static particles never move
            assert (out[name][i] == np.asarray(pd.getPosition0(i), dtype=np.float32)).all()
        assert out[name][:, 1].min() < -1e-3               # the rest sags under gravity
    assert pbd.Timing.averageStepMs() > 0.0
