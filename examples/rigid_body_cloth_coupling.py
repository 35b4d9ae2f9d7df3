"""Headless counterpart of the reference's pyPBD/examples/rigid_body_cloth_coupling.py: a 20 x 20 cloth hung between four
three-body chains (static anchor + two dynamic boxes connected by ball joints), its corners attached to the last body of each
chain by rigid-body-particle ball joints (Demos/CouplingDemos/RigidBodyClothCouplingDemo.cpp:151-289).  The reference loads
data/models/cube.obj; the unit cube is generated here instead (mesh import is outside this path).  Needs a CUDA device."""
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import positionbaseddynamics_b200.pypbd as pbd

nRows, nCols = 20, 20
clothWidth, clothHeight = 10.0, 10.0
width, height, depth = 0.4, 2.0, 0.4     # the dynamic boxes

CUBE_V = np.array([[-0.5, -0.5, -0.5], [0.5, -0.5, -0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, -0.5],
                   [-0.5, -0.5, 0.5], [0.5, -0.5, 0.5], [0.5, 0.5, 0.5], [-0.5, 0.5, 0.5]])
CUBE_F = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [1, 2, 6], [1, 6, 5], [0, 4, 7], [0, 7, 3]])


def rotation_x(angle):
    c, s = math.cos(angle), math.sin(angle)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def buildModel(simModel=2, bendingModel=2):
    sim = pbd.Simulation.getCurrent()
    sim.initDefault()
    model = sim.getModel()
    triModel = model.addRegularTriangleModel(nCols, nRows, [-5, 4, -5], rotation_x(math.pi * 0.5), [clothWidth, clothHeight], testMesh=False)
    stiffness = 100000 if simModel == 4 else 1.0
    model.addClothConstraints(triModel, simModel, stiffness, stiffness, stiffness, stiffness, 0.3, 0.3, False, False)
    model.addBendingConstraints(triModel, bendingModel, 50.0 if bendingModel == 3 else 0.01)
    for (cx, cz) in ((-5.0, -5.0), (5.0, -5.0), (5.0, 5.0), (-5.0, 5.0)):
        anchor = model.addRigidBody(1.0, CUBE_V, CUBE_F, translation=[cx, 0.0, cz], scale=[0.5, 0.5, 0.5], testMesh=False, generateCollisionObject=False)
        anchor.setMass(0.0)
        model.addRigidBody(1.0, CUBE_V, CUBE_F, [cx, 1.0, cz], scale=[width, height, depth], testMesh=False, generateCollisionObject=False)
        model.addRigidBody(1.0, CUBE_V, CUBE_F, [cx, 3.0, cz], scale=[width, height, depth], testMesh=False, generateCollisionObject=False)
    for chain in range(4):
        base = 3 * chain
        x, z = model.getRigidBodies()[base].getPosition()[[0, 2]]
        model.addBallJoint(base, base + 1, [x, 0.0, z])
        model.addBallJoint(base + 1, base + 2, [x, 2.0, z])
    model.addRigidBodyParticleBallJoint(2, 0)
    model.addRigidBodyParticleBallJoint(5, nCols - 1)
    model.addRigidBodyParticleBallJoint(8, nRows * nCols - 1)
    model.addRigidBodyParticleBallJoint(11, (nRows - 1) * nCols)
    sim.getTimeStep().setValueUInt(pbd.TimeStepController.NUM_SUB_STEPS, 3)
    return model


def main(frames=10):
    pbd.Logger.addConsoleSink(pbd.LogLevel.INFO)
    model = buildModel()
    sim = pbd.Simulation.getCurrent()
    for _ in range(frames):
        for _ in range(8):
            sim.getTimeStep().step(model)
    x = model.getParticles().getVertices()
    print("Time: {:.2f}".format(pbd.TimeManager.getCurrent().getTime()), "cloth centroid", x.mean(axis=0))
    for i, rb in enumerate(model.getRigidBodies()):
        print("body %2d mass %.3f position" % (i, rb.getMass()), np.round(rb.getPosition(), 4))
    return x


if __name__ == "__main__":
    main()
