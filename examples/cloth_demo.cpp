// examples/cloth_demo.cpp -- headless C++ counterpart of the reference's Demos/ClothDemo/main.cpp (buildModel / createMesh /
// timeStep, :119-162, :100-117) written against the host-side C++ mirror of the reference interface
// (positionbaseddynamics_b200/csrc/host/pbd_model.h: same class and method names as Simulation/SimulationModel.h,
// ParticleData.h, TimeStepController.h).  Scene construction is line for line what the demo does; stepping goes to the GPU.
//
//   g++ -std=c++17 -I. examples/cloth_demo.cpp -Lpositionbaseddynamics_b200 -lpbd_b200 -Wl,-rpath,$PWD/positionbaseddynamics_b200 -o cloth_demo
//
// Without a CUDA device the time step reports "no CUDA device available ... no CPU fallback" and the program exits with 2.
#include <cmath>
#include <cstdio>
#include "positionbaseddynamics_b200/csrc/host/pbd_model.h"

using namespace pbd_b200;

int main(int argc, char **argv) {
    const int nRows = 50, nCols = 50;
    const Real width = 10.0f, height = 10.0f;
    const int steps = argc > 1 ? atoi(argv[1]) : 200;
    const unsigned int simulationMethod = 2, bendingMethod = 2;  // ClothDemo defaults: FEM triangles + isometric bending

    SimulationModel model;
    model.init();

    // createMesh(): a regular grid in the x-z plane (the demo rotates the x-y grid by 90 degrees about x)
    Matrix3r R = Matrix3r::Identity();
    const Real a = static_cast<Real>(M_PI * 0.5);
    R(1, 1) = std::cos(a); R(1, 2) = -std::sin(a); R(2, 1) = std::sin(a); R(2, 2) = std::cos(a);
    model.addRegularTriangleModel(nCols, nRows, Vector3r(0, 1, 0), R, Vector2r{{width, height}});
    ParticleData &pd = model.getParticles();
    for (unsigned int i = 0; i < pd.getNumberOfParticles(); i++) pd.setMass(i, 1.0f);
    pd.setMass(0, 0.0f);             // two corners are static
    pd.setMass(nRows - 1, 0.0f);     // (the demo pins particle (nRows-1)*nCols; pyPBD's cloth_model.py pins nRows-1: same corner row)

    for (TriangleModel *tm : model.getTriangleModels()) {
        model.addClothConstraints(tm, simulationMethod, 1.0f, 1.0f, 1.0f, 1.0f, 0.3f, 0.3f, false, false);
        model.addBendingConstraints(tm, bendingMethod, 0.01f);
    }
    std::printf("Number of triangles: %u\nNumber of vertices: %d\nNumber of constraints: %u\n",
                model.getTriangleModels()[0]->getParticleMesh().numFaces(), nRows * nCols, model.numConstraints());

    TimeStepController ts;   // owns the GPU engine; the reference obtains its TimeStepController from Simulation::getCurrent()
    ts.init();
    if (!ts.valid()) { std::fprintf(stderr, "TimeStepController: %s\n", ts.error().c_str()); return 2; }
    ts.timeManager().setTimeStepSize(0.005f);
    ts.setValueUInt(TimeStepController::NUM_SUB_STEPS, 5);
    ts.setValueUInt(TimeStepController::MAX_ITERATIONS, 1);

    for (int i = 0; i < steps; i++)
        if (!ts.step(model)) { std::fprintf(stderr, "step failed: %s\n", ts.error().c_str()); return 1; }

    const ParticleData &cpd = model.getParticles();  // const accessors pull the state back from the device
    double c[3] = {0, 0, 0};
    for (unsigned int i = 0; i < cpd.size(); i++) for (int k = 0; k < 3; k++) c[k] += cpd.getPosition(i)[k];
    std::printf("Time: %.3f centroid %.5f %.5f %.5f\n", (double)ts.timeManager().getTime(), c[0] / cpd.size(), c[1] / cpd.size(), c[2] / cpd.size());
    return 0;
}
