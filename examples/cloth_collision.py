"""Headless counterpart of the reference's Demos/DistanceFieldDemos/ClothCollisionDemo.cpp (the pyPBD example cloth_collision.py uses
CubicSDFCollisionDetection, whose grid library is not vendored): a 50 x 50 XPBD cloth falls onto a static torus lying on a static floor
box.  Same calls as the demo -- addRigidBody + setMass(0), DistanceFieldCollisionDetection.addCollisionBox / addCollisionTorus,
addCollisionObjectWithoutGeometry for the cloth, TimeStep.setCollisionDetection -- through positionbaseddynamics_b200.pypbd.  Collision
test and velocity-level contact solve run on the GPU (include/pbd_b200.h, "Contact path").  Needs a CUDA device."""
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import positionbaseddynamics_b200.pypbd as pbd

CUBE_V = np.array([[-0.5, -0.5, -0.5], [0.5, -0.5, -0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, -0.5],
                   [-0.5, -0.5, 0.5], [0.5, -0.5, 0.5], [0.5, 0.5, 0.5], [-0.5, 0.5, 0.5]])
CUBE_F = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [1, 2, 6], [1, 6, 5], [0, 4, 7], [0, 7, 3]])
nRows, nCols = 50, 50
width, height = 10.0, 10.0


def buildModel():
    sim = pbd.Simulation.getCurrent()
    sim.initDefault()
    model = sim.getModel()
    a = math.pi * 0.5
    R = [[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]]
    triModel = model.addRegularTriangleModel(nCols, nRows, [-5, 4, -5], R, [width, height], testMesh=False)   # ClothCollisionDemo.cpp:200-203
    model.addClothConstraints(triModel, 4, 1.0e5, 1.0, 1.0, 1.0, 0.3, 0.3, False, False)
    model.addBendingConstraints(triModel, 3, 100.0)
    # static floor and torus (ClothCollisionDemo.cpp:134-160): the meshes only feed the bounding boxes and the mass properties
    floor = model.addRigidBody(1.0, CUBE_V, CUBE_F, [0.0, -0.5, 0.0], np.eye(3), [100.0, 1.0, 100.0])
    floor.setMass(0.0)
    torus = model.addRigidBody(1.0, CUBE_V, CUBE_F, [0.0, 1.5, 0.0], np.eye(3), [6.0, 2.0, 6.0])
    torus.setMass(0.0); torus.setFrictionCoeff(0.1)
    cd = pbd.DistanceFieldCollisionDetection()
    sim.getTimeStep().setCollisionDetection(model, cd)
    cd.setTolerance(0.05)
    T = pbd.CollisionObject
    cd.addCollisionBox(0, T.RigidBodyCollisionObjectType, CUBE_V * [100.0, 1.0, 100.0], 8, [100.0, 1.0, 100.0])
    cd.addCollisionTorus(1, T.RigidBodyCollisionObjectType, CUBE_V * [6.0, 2.0, 6.0], 8, [2.0, 1.0])
    for i, tm in enumerate(model.getTriangleModels()):
        tm.setFrictionCoeff(0.1)
        cd.addCollisionObjectWithoutGeometry(i, T.TriangleModelCollisionObjectType, None, 0, True)
    ts = sim.getTimeStep()
    ts.setValueUInt(pbd.TimeStepController.NUM_SUB_STEPS, 1)
    ts.setValueUInt(pbd.TimeStepController.MAX_ITERATIONS, 5)
    return model


def main(frames=60):
    model = buildModel()
    sim = pbd.Simulation.getCurrent()
    for _ in range(frames):
        for _ in range(8):
            sim.getTimeStep().step(model)
    x = model.getParticles().getVertices()
    print("Time: {:.2f}".format(pbd.TimeManager.getCurrent().getTime()), "lowest particle y = %.3f" % x[:, 1].min(), "(floor at 0, torus top at 2.5)")
    return x


if __name__ == "__main__":
    main()
