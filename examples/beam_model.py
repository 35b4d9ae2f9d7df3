"""Headless counterpart of the reference's pyPBD/examples/beam_model.py: a 30 x 5 x 5 tetrahedral bar clamped at one end
(solid model 2 = FEM tets by default), driven through positionbaseddynamics_b200.pypbd.  Prints the tip position instead of
rendering.  Needs a CUDA device."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import positionbaseddynamics_b200.pypbd as pbd

width, depth, height = 30, 5, 5


def buildModel(simModel=2):
    sim = pbd.Simulation.getCurrent()
    sim.initDefault()
    model = sim.getModel()
    tetModel = model.addRegularTetModel(width, height, depth, [5, 0, 0], np.eye(3), [10.0, 1.5, 1.5], testMesh=False)
    pd = model.getParticles()
    for i in range(1):          # clamp the x = 0 slab, as the reference example does
        for j in range(height):
            for k in range(depth):
                pd.setMass(i * height * depth + j * depth + k, 0.0)
    stiffness = 100000 if simModel in (3, 6) else 1.0
    volumeStiffness = 100000 if simModel == 6 else 1.0
    model.addSolidConstraints(tetModel, simModel, stiffness, 0.3, volumeStiffness, False, False)
    print("Number of tets: " + str(tetModel.getParticleMesh().numTets()))
    print("Number of vertices: " + str(width * height * depth))
    return model


def main(frames=20, simModel=2):
    pbd.Logger.addConsoleSink(pbd.LogLevel.INFO)
    model = buildModel(simModel)
    sim = pbd.Simulation.getCurrent()
    for _ in range(frames):
        for _ in range(8):
            sim.getTimeStep().step(model)
    x = model.getParticles().getVertices()
    print("Time: {:.2f}".format(pbd.TimeManager.getCurrent().getTime()), "tip", x[-1])
    return x


if __name__ == "__main__":
    main()
