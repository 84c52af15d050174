// The glue's integration tests on the (shim) CUDA platform: B200CudaIntegrateRigidBodyStepKernel works on the
// CudaContext's own device arrays (posq [+ posqCorrection], velm, fixed-point force planes, reordered atoms) - the
// counterpart of the reference's platforms/cuda/tests/TestCudaRigidBodyIntegrator.cpp.
//   TestB200CudaRigidBodyIntegrator [dump-file]
// With a file name, the last test also writes its inputs and results there for the Python-side comparison with the CPU
// oracle (tests/test_glue.py).
#include "RigidBodyTests.h"
#include "rbk.h"
#include <fstream>

static void buildWaters(int nMol, unsigned seed0, System& system, vector<int>& bodyIndices, vector<Vec3>& positions, vector<Vec3>& velocities,
                        vector<double>& charges) {
    const double rOH = 0.09572, half = 0.5*104.52*M_PI/180.0;
    unsigned seed = seed0;
    auto rnd = [&]() { seed = seed*1664525u + 1013904223u; return (seed >> 8)/16777216.0 - 0.5; };
    const int side = (int) ceil(cbrt((double) nMol));
    for (int m = 0; m < nMol; m++) {
        Vec3 c(0.31*(m % side), 0.31*((m/side) % side), 0.31*(m/(side*side)));
        // a random orientation: two successive rotations of the template about z and x
        const double a = 6.0*rnd(), b = 6.0*rnd();
        Vec3 site[3] = {Vec3(0, 0, 0), Vec3(rOH*sin(half), 0, rOH*cos(half)), Vec3(-rOH*sin(half), 0, rOH*cos(half))};
        const double mass[3] = {15.99943, 1.007947, 1.007947}, q[3] = {-0.834, 0.417, 0.417};
        for (int k = 0; k < 3; k++) {
            Vec3 x = site[k];
            x = Vec3(cos(a)*x[0] - sin(a)*x[1], sin(a)*x[0] + cos(a)*x[1], x[2]);
            x = Vec3(x[0], cos(b)*x[1] - sin(b)*x[2], sin(b)*x[1] + cos(b)*x[2]);
            system.addParticle(mass[k]);
            positions.push_back(c + x);
            velocities.push_back(Vec3(rnd(), rnd(), rnd())*(k == 0 ? 0.7 : 2.7));       // about 300 K
            bodyIndices.push_back(m + 1);
            charges.push_back(q[k]);
        }
    }
}

// Tethered waters (every atom feels the analytic test potential, so bodies get forces AND torques), stepped on a platform.
struct Run {
    vector<Vec3> positions, velocities;
    vector<double> ke;
    double pe;
    int reorders;
    long long copyCalls;
    int stateUpdates;
};
static Run runTethered(Platform& platform, int nMol, int mode, int steps, int chunk, int reorderInterval, const string& dump = "",
                       int reorderUnit = 1) {
    System system;
    vector<int> bodyIndices;
    vector<Vec3> positions, velocities;
    vector<double> charges;
    buildWaters(nMol, 2468u, system, bodyIndices, positions, velocities, charges);
    const double k = 5000.0;
    const Vec3 field(300.0, -500.0, 800.0);
    CustomExternalForce* tether = new CustomExternalForce(k, field);
    for (size_t i = 0; i < positions.size(); i++) tether->addParticle((int) i, positions[i], charges[i]);
    system.addForce(tether);
    RigidBodyIntegrator integrator(0.001, bodyIndices);
    integrator.setRotationMode(mode);
    Context context(system, integrator, platform);
    CudaContext* cu = NULL;
    if (platform.getName() == "CUDA") {
        // reach the CudaContext the way the kernel factory does
        cu = static_cast<CudaPlatform::PlatformData*>(integrator.getContextImpl().getPlatformData())->contexts[0];
        cu->setReorderInterval(reorderInterval);
        cu->setReorderUnit(reorderUnit);
    }
    context.setPositions(positions);
    context.setVelocities(velocities);
    long long before[4], after[4];
    rbk_debug_copy_counters(integrator.getRigidBodySystem().getHandle(), before);
    const int updates0 = integrator.getContextImpl().getNumStateUpdates();
    for (int done = 0; done < steps; done += chunk) integrator.step(min(chunk, steps - done));
    rbk_debug_copy_counters(integrator.getRigidBodySystem().getHandle(), after);
    Run r;
    State s = context.getState(State::Positions | State::Velocities | State::Energy);
    r.positions = s.getPositions();
    r.velocities = s.getVelocities();
    r.pe = s.getPotentialEnergy();
    r.ke = integrator.getKineticEnergies();
    r.reorders = cu ? cu->getNumReorders() : 0;
    r.copyCalls = (after[0] - before[0]) + (after[2] - before[2]);
    r.stateUpdates = integrator.getContextImpl().getNumStateUpdates() - updates0;
    ASSERT_TOL(0.001*steps, context.getState(0).getTime(), 1e-9);
    if (!dump.empty()) {
        ofstream out(dump.c_str(), ios::binary);
        const int n = (int) positions.size();
        const double header[8] = {(double) n, (double) nMol, (double) mode, (double) steps, 0.001, k, r.ke[0], r.ke[1]};
        out.write((const char*) header, sizeof(header));
        out.write((const char*) &field, sizeof(Vec3));
        out.write((const char*) charges.data(), n*sizeof(double));
        out.write((const char*) positions.data(), n*sizeof(Vec3));
        out.write((const char*) velocities.data(), n*sizeof(Vec3));
        out.write((const char*) r.positions.data(), n*sizeof(Vec3));
        out.write((const char*) r.velocities.data(), n*sizeof(Vec3));
    }
    return r;
}

static double maxDiff(const vector<Vec3>& a, const vector<Vec3>& b) {
    double worst = 0.0, scale = 0.0;
    for (size_t i = 0; i < a.size(); i++)
        for (int c = 0; c < 3; c++) { worst = max(worst, fabs(a[i][c] - b[i][c])); scale = max(scale, fabs(b[i][c])); }
    return worst/scale;
}

// The same physical system through both kernels of the plugin: host vectors on the Reference platform (itself checked
// against the oracle by the Python suite) and device arrays on the CUDA platform - fused step(n), one step at a time,
// with and without atom reordering.  Reordering must not change a single bit; fusing only the summation order of a
// water's three atom forces.
static void testCudaMatchesReferencePlatform(Platform& reference, Platform& cuda, int mode, const string& dump) {
    const int nMol = 700, steps = 24;
    Run ref = runTethered(reference, nMol, mode, steps, steps, 0);
    Run fused = runTethered(cuda, nMol, mode, steps, steps, 0);
    Run fusedReordered = runTethered(cuda, nMol, mode, steps, 8, 3, dump);
    Run single = runTethered(cuda, nMol, mode, steps, 1, 5);
    Run molecules = runTethered(cuda, nMol, mode, steps, 6, 2, "", 3);       // whole waters move: the handle re-sorts its bodies
    ASSERT(fusedReordered.reorders >= 7 && single.reorders >= 4 && fused.reorders == 0 && molecules.reorders >= 11);
    ASSERT(maxDiff(molecules.positions, fused.positions) < 1e-13 && maxDiff(molecules.velocities, fused.velocities) < 1e-12);
    ASSERT(maxDiff(fused.positions, ref.positions) < 1e-11 && maxDiff(fused.velocities, ref.velocities) < 1e-10);
    ASSERT(maxDiff(single.positions, ref.positions) < 1e-11 && maxDiff(single.velocities, ref.velocities) < 1e-10);
    ASSERT(maxDiff(fusedReordered.positions, fused.positions) < 1e-13 && maxDiff(fusedReordered.velocities, fused.velocities) < 1e-12);
    ASSERT_TOL(ref.ke[0] + ref.ke[1], fused.ke[0] + fused.ke[1], 1e-10);
    ASSERT_TOL(ref.pe, fused.pe, 1e-9);
    // nothing librbk does per step moves data between host and device: zero copies issued by the library in step(n)
    ASSERT(fused.copyCalls == 0);
    // reorders go through rbk_reorder_openmm: a few small uploads (the new location table, re-sort indices) each, nothing else
    ASSERT(fusedReordered.copyCalls >= fusedReordered.reorders && fusedReordered.copyCalls <= 4*fusedReordered.reorders);
    // updateContextState once per step, like CudaRigidBodyKernels.cpp:378
    ASSERT(fused.stateUpdates == steps && single.stateUpdates == steps);
}

int main(int argc, char** argv) {
    try {
        const string dump = argc > 1 ? argv[1] : "";
        ReferencePlatform* reference = new ReferencePlatform();
        Platform::registerPlatform(reference);
        registerRigidBodyCudaKernelFactories();               // registers the (shim) CUDA platform if there is none yet
        Platform& cuda = Platform::getPlatformByName("CUDA");
        const char* precisions[] = {"mixed", "double"};
        for (int p = 0; p < 2; p++) {
            cuda.setPropertyDefaultValue(CudaPlatform::CudaPrecision(), precisions[p]);
            testErrors(cuda);
            testSingleBond(cuda);
            testRigidWaters(cuda, 0);
            testRigidWaters(cuda, 3);
            // the CUDA flow constrains the displacement before the move (CudaRigidBodyKernels.cpp:405-421); what the solver
            // did reaches the velocities in Part 2, as on the Reference platform
            testConstrainedFreeAtoms(cuda);
            testRefinedEnergies(cuda);
            testCudaMatchesReferencePlatform(*reference, cuda, 0, p == 0 ? dump : "");
            testCudaMatchesReferencePlatform(*reference, cuda, 4, "");
        }
    }
    catch (const exception& e) {
        cout << "exception: " << e.what() << endl;
        return 1;
    }
    cout << "Done" << endl;
    return 0;
}
