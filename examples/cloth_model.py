"""Headless counterpart of the reference's pyPBD/examples/cloth_model.py: the same model-building calls (50 x 50 cloth hanging
from two corners, cloth model 2 = FEM triangles, bending model 2 = isometric bending, 3 substeps), driven through
positionbaseddynamics_b200.pypbd instead of pypbd.  Rendering (pygame / OpenGL) is outside this path, so the script prints the
centroid instead.  Needs a CUDA device: the engine has no CPU fallback."""
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import positionbaseddynamics_b200.pypbd as pbd

nRows, nCols = 50, 50
width, height = 10.0, 10.0


def rotation_matrix(angle, axis):
    x, y, z = np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
    c, s = math.cos(angle), math.sin(angle)
    return np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                     [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                     [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])


def buildModel(simModel=2, bendingModel=2):
    sim = pbd.Simulation.getCurrent()
    sim.initDefault()
    model = sim.getModel()
    triModel = model.addRegularTriangleModel(nCols, nRows, [0, 0, 0], rotation_matrix(math.pi * 0.5, [1.0, 0.0, 0.0]), [width, height], testMesh=False)
    pd = model.getParticles()
    pd.setMass(0, 0.0)
    pd.setMass(nRows - 1, 0.0)
    stiffness = 100000 if simModel == 4 else 1.0
    model.addClothConstraints(triModel, simModel, stiffness, stiffness, stiffness, stiffness, 0.3, 0.3, False, False)
    model.addBendingConstraints(triModel, bendingModel, 50.0 if bendingModel == 3 else 0.01)
    print("Number of triangles: " + str(triModel.getParticleMesh().numFaces()))
    print("Number of vertices: " + str(nRows * nCols))
    sim.getTimeStep().setValueUInt(pbd.TimeStepController.NUM_SUB_STEPS, 3)
    return model


def main(frames=25, simModel=2, bendingModel=2):
    pbd.Logger.addConsoleSink(pbd.LogLevel.INFO)
    pbd.Timing.reset(); pbd.Timing.enabled = True
    model = buildModel(simModel, bendingModel)
    sim = pbd.Simulation.getCurrent()
    for _ in range(frames):
        for _ in range(8):  # the reference's timeStep() advances 8 steps per rendered frame
            sim.getTimeStep().step(model)
    x = model.getParticles().getVertices()
    print("Time: {:.2f}".format(pbd.TimeManager.getCurrent().getTime()), "centroid", x.mean(axis=0))
    pbd.Timing.printAverageTimes()
    return x


if __name__ == "__main__":
    main()
