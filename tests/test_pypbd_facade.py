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
