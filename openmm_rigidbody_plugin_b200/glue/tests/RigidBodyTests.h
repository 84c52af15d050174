// Integration tests of the C++ glue on the GPU, shared by the two platform test programs the way the reference shares
// tests/TestRigidBodyIntegrator.h between its Reference and CUDA tests: testSingleBond is the reference's own test (two
// free atoms on a harmonic bond against the analytic solution, tests/TestRigidBodyIntegrator.h:49-83); the other tests
// exercise what that file leaves commented out or never covered: actual rigid bodies.
#ifndef RBK_GLUE_RIGIDBODYTESTS_H_
#define RBK_GLUE_RIGIDBODYTESTS_H_
#include "B200RigidBodyKernelFactory.h"
#include "RigidBodyIntegrator.h"
#include "openmm/Context.h"
#include "openmm/CustomExternalForce.h"
#include "openmm/HarmonicBondForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/VirtualSite.h"
#include "openmm/cuda/CudaPlatform.h"
#include "openmm/reference/ReferencePlatform.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

using namespace RigidBodyPlugin;
using namespace OpenMM;
using namespace std;

#define ASSERT(cond) do { if (!(cond)) throw runtime_error(string("assertion failed: ") + #cond + " (line " + to_string(__LINE__) + ")"); } while (0)
#define ASSERT_TOL(expected, found, tol) do { double e_ = (expected), f_ = (found); \
    if (!(fabs(e_ - f_) <= (tol)*max(1.0, fabs(e_)))) throw runtime_error("expected " + to_string(e_) + " found " + to_string(f_) + " (line " + to_string(__LINE__) + ")"); } while (0)
#define ASSERT_VEC(expected, found, tol) do { Vec3 e_ = (expected), f_ = (found); \
    for (int c_ = 0; c_ < 3; c_++) if (!(fabs(e_[c_] - f_[c_]) <= (tol))) throw runtime_error("vector mismatch (line " + to_string(__LINE__) + ")"); } while (0)

static void testSingleBond(Platform& platform) {
    System system;
    system.addParticle(2.0);
    system.addParticle(2.0);
    vector<int> bodyIndices(2, 0);
    RigidBodyIntegrator integrator(0.01, bodyIndices);
    HarmonicBondForce* bond = new HarmonicBondForce();
    bond->addBond(0, 1, 1.5, 1);
    system.addForce(bond);
    Context context(system, integrator, platform);
    vector<Vec3> positions(2);
    positions[0] = Vec3(-1, 0, 0);
    positions[1] = Vec3(1, 0, 0);
    context.setPositions(positions);
    const double freq = 1.0;
    State state = context.getState(State::Energy);
    const double initialEnergy = state.getKineticEnergy() + state.getPotentialEnergy();
    for (int i = 0; i < 1000; ++i) {
        state = context.getState(State::Positions | State::Velocities | State::Energy);
        double time = state.getTime();
        double dist = 1.5 + 0.5*cos(freq*time);
        ASSERT_VEC(Vec3(-0.5*dist, 0, 0), state.getPositions()[0], 0.02);
        ASSERT_VEC(Vec3(0.5*dist, 0, 0), state.getPositions()[1], 0.02);
        double speed = -0.5*freq*sin(freq*time);
        ASSERT_VEC(Vec3(-0.5*speed, 0, 0), state.getVelocities()[0], 0.02);
        ASSERT_VEC(Vec3(0.5*speed, 0, 0), state.getVelocities()[1], 0.02);
        ASSERT_TOL(initialEnergy, state.getKineticEnergy() + state.getPotentialEnergy(), 0.01);
        integrator.step(1);
    }
    ASSERT_TOL(10.0, context.getState(0).getTime(), 1e-5);
}

// 64 rigid waters coupled by harmonic bonds between neighbouring oxygens: bodies stay rigid, total energy is
// conserved to the integrator's accuracy, DOF and the kinetic-energy split are reported like the reference's API.
static void testRigidWaters(Platform& platform, int mode) {
    const int nMol = 64;
    const double rOH = 0.09572, half = 0.5*104.52*M_PI/180.0;
    System system;
    vector<int> bodyIndices;
    vector<Vec3> positions, velocities;
    unsigned seed = 12345u;
    auto rnd = [&]() { seed = seed*1664525u + 1013904223u; return (seed >> 8)/16777216.0 - 0.5; };
    for (int m = 0; m < nMol; m++) {
        Vec3 c(0.35*(m % 4), 0.35*((m/4) % 4), 0.35*(m/16));
        system.addParticle(15.99943); system.addParticle(1.007947); system.addParticle(1.007947);
        positions.push_back(c);
        positions.push_back(c + Vec3(rOH*sin(half), 0, rOH*cos(half)));
        positions.push_back(c + Vec3(-rOH*sin(half), 0, rOH*cos(half)));
        for (int k = 0; k < 3; k++) { bodyIndices.push_back(m + 1); velocities.push_back(Vec3(rnd(), rnd(), rnd())); }
    }
    HarmonicBondForce* bonds = new HarmonicBondForce();
    for (int m = 0; m + 1 < nMol; m++) bonds->addBond(3*m, 3*(m + 1), 0.33, 2000.0);
    for (int m = 0; m + 4 < nMol; m++) bonds->addBond(3*m + 1, 3*(m + 4) + 2, 0.36, 500.0);
    system.addForce(bonds);
    RigidBodyIntegrator integrator(0.001, bodyIndices);
    integrator.setRotationMode(mode);
    Context context(system, integrator, platform);
    context.setPositions(positions);
    context.setVelocities(velocities);
    ASSERT(integrator.getRigidBodySystem().getNumBodies() == nMol);
    ASSERT(integrator.getRigidBodySystem().getNumFree() == 0);
    ASSERT(integrator.getRigidBodySystem().getNumDOF() == 6*nMol);
    State s0 = context.getState(State::Energy);
    const double e0 = s0.getKineticEnergy() + s0.getPotentialEnergy();
    for (int block = 0; block < 10; block++) {
        integrator.step(100);
        State s = context.getState(State::Positions | State::Velocities | State::Energy);
        ASSERT_TOL(e0, s.getKineticEnergy() + s.getPotentialEnergy(), 2e-3);
        vector<double> ke = integrator.getKineticEnergies();
        ASSERT_TOL(s.getKineticEnergy(), ke[0] + ke[1], 1e-12);
        ASSERT_TOL(ke[0] + ke[1], integrator.getRigidBodySystem().getKineticEnergy(), 1e-12);
        for (int m = 0; m < nMol; m++) {                       // every molecule is still a rigid TIP3P water
            Vec3 a = s.getPositions()[3*m + 1] - s.getPositions()[3*m], b = s.getPositions()[3*m + 2] - s.getPositions()[3*m];
            ASSERT_TOL(rOH, sqrt(a.dot(a)), 1e-11);
            ASSERT_TOL(rOH, sqrt(b.dot(b)), 1e-11);
            ASSERT_TOL(cos(2*half), a.dot(b)/(rOH*rOH), 1e-10);
            Vec3 dv = s.getVelocities()[3*m + 1] - s.getVelocities()[3*m];
            ASSERT(fabs(dv.dot(a)) < 1e-10);                   // no velocity along a rigid bond
        }
    }
    ASSERT_TOL(1.0, context.getState(0).getTime(), 1e-9);
    vector<double> refined = integrator.getRefinedKineticEnergies();
    vector<double> plain = integrator.getKineticEnergies();
    ASSERT(refined[0] == plain[0] && refined[1] == plain[1]);
    ASSERT(integrator.getPotentialEnergyRefinement() == 0.0);
}

// Rigid waters + constrained free diatomics (+ a virtual site at one bond's midpoint): the path on which the reference's
// kernel calls ReferenceConstraints::apply / applyToVelocities and ReferenceVirtualSites::computePositions
// (ReferenceRigidBodyKernels.cpp:92-104).  Constraints hold to the integrator's tolerance, the velocity along each
// constrained bond vanishes, energy is conserved, DOF = numFree - numConstraints + 6 per body (RigidBodySystem.cpp:130-134).
static void testConstrainedFreeAtoms(Platform& platform, double energyTol = 2e-3) {
    const int nMol = 16, nPairs = 12;
    const double rOH = 0.09572, half = 0.5*104.52*M_PI/180.0, bond = 0.12;
    System system;
    vector<int> bodyIndices;
    vector<Vec3> positions, velocities;
    unsigned seed = 777u;
    auto rnd = [&]() { seed = seed*1664525u + 1013904223u; return (seed >> 8)/16777216.0 - 0.5; };
    for (int m = 0; m < nMol; m++) {
        Vec3 c(0.35*(m % 4), 0.35*(m/4), 0.0);
        system.addParticle(15.99943); system.addParticle(1.007947); system.addParticle(1.007947);
        positions.push_back(c);
        positions.push_back(c + Vec3(rOH*sin(half), 0, rOH*cos(half)));
        positions.push_back(c + Vec3(-rOH*sin(half), 0, rOH*cos(half)));
        for (int k = 0; k < 3; k++) { bodyIndices.push_back(m + 1); velocities.push_back(Vec3(rnd(), rnd(), rnd())); }
    }
    const int firstFree = 3*nMol;
    HarmonicBondForce* bonds = new HarmonicBondForce();
    for (int k = 0; k < nPairs; k++) {
        Vec3 c(0.35*(k % 4) + 0.1, 0.35*(k/4) + 0.1, 0.4);
        const int a = system.addParticle(14.0), b = system.addParticle(16.0);
        positions.push_back(c);
        positions.push_back(c + Vec3(bond, 0, 0));
        Vec3 v(rnd(), rnd(), rnd()), w(0, rnd(), rnd());            // relative velocity perpendicular to the bond
        velocities.push_back(v);
        velocities.push_back(v + w);
        bodyIndices.push_back(0); bodyIndices.push_back(0);
        system.addConstraint(a, b, bond);
        bonds->addBond(a, 3*k, 0.42, 800.0);                        // tie the diatomics to water oxygens
        bonds->addBond(b, 3*((k + 5) % nMol) + 1, 0.45, 300.0);
    }
    const int site = system.addParticle(0.0);
    system.setVirtualSite(site, new TwoParticleAverageSite(firstFree, firstFree + 1, 0.25, 0.75));
    positions.push_back(positions[firstFree]*0.25 + positions[firstFree + 1]*0.75);
    velocities.push_back(Vec3());
    bodyIndices.push_back(0);
    bonds->addBond(site, 0, 0.5, 100.0);                            // felt by the two particles that define the site
    system.addForce(bonds);
    RigidBodyIntegrator integrator(0.001, bodyIndices);
    integrator.setConstraintTolerance(1e-9);
    Context context(system, integrator, platform);
    context.setPositions(positions);
    context.setVelocities(velocities);
    ASSERT(integrator.getRigidBodySystem().getNumBodies() == nMol);
    ASSERT(integrator.getRigidBodySystem().getNumFree() == 2*nPairs);
    ASSERT(integrator.getRigidBodySystem().getNumDOF() == 2*nPairs - nPairs + 6*nMol);
    State s0 = context.getState(State::Energy);
    const double e0 = s0.getKineticEnergy() + s0.getPotentialEnergy();
    double moved = 0.0;
    for (int block = 0; block < 10; block++) {
        integrator.step(50);
        State s = context.getState(State::Positions | State::Velocities | State::Energy);
        const vector<Vec3>& R = s.getPositions();
        const vector<Vec3>& V = s.getVelocities();
        for (int k = 0; k < nPairs; k++) {
            Vec3 r = R[firstFree + 2*k] - R[firstFree + 2*k + 1], v = V[firstFree + 2*k] - V[firstFree + 2*k + 1];
            ASSERT_TOL(bond, sqrt(r.dot(r)), 1e-8);
            ASSERT(fabs(r.dot(v)) < 1e-8);
        }
        ASSERT_VEC(R[firstFree]*0.25 + R[firstFree + 1]*0.75, R[site], 1e-14);
        Vec3 a = R[1] - R[0];
        ASSERT_TOL(rOH, sqrt(a.dot(a)), 1e-11);
        ASSERT_TOL(e0, s.getKineticEnergy() + s.getPotentialEnergy(), energyTol);
        Vec3 d = R[firstFree] - positions[firstFree];
        moved = max(moved, sqrt(d.dot(d)));
    }
    ASSERT(moved > 1e-3);
}

// setComputeRefinedEnergies(true): the refined kinetic energies + potential refinement (the reference's CUDA-only
// diagnostics, CudaRigidBodyKernels.cpp:469-494) make a total energy that fluctuates far less than the plain one.
static void testRefinedEnergies(Platform& platform) {
    const int nMol = 27;
    const double rOH = 0.09572, half = 0.5*104.52*M_PI/180.0;
    System system;
    vector<int> bodyIndices;
    vector<Vec3> positions, velocities;
    unsigned seed = 4321u;
    auto rnd = [&]() { seed = seed*1664525u + 1013904223u; return (seed >> 8)/16777216.0 - 0.5; };
    for (int m = 0; m < nMol; m++) {
        Vec3 c(0.35*(m % 3), 0.35*((m/3) % 3), 0.35*(m/9));
        system.addParticle(15.99943); system.addParticle(1.007947); system.addParticle(1.007947);
        positions.push_back(c);
        positions.push_back(c + Vec3(rOH*sin(half), 0, rOH*cos(half)));
        positions.push_back(c + Vec3(-rOH*sin(half), 0, rOH*cos(half)));
        for (int k = 0; k < 3; k++) { bodyIndices.push_back(m + 1); velocities.push_back(Vec3(rnd(), rnd(), rnd())); }
    }
    HarmonicBondForce* bonds = new HarmonicBondForce();
    for (int m = 0; m + 1 < nMol; m++) bonds->addBond(3*m, 3*(m + 1), 0.33, 2000.0);
    for (int m = 0; m + 3 < nMol; m++) bonds->addBond(3*m + 1, 3*(m + 3) + 2, 0.36, 500.0);
    system.addForce(bonds);
    RigidBodyIntegrator integrator(0.001, bodyIndices);
    integrator.setComputeRefinedEnergies(true);
    Context context(system, integrator, platform);
    context.setPositions(positions);
    context.setVelocities(velocities);
    vector<double> k0 = integrator.getKineticEnergies(), r0 = integrator.getRefinedKineticEnergies();
    ASSERT(k0[0] == r0[0] && k0[1] == r0[1]);                      // nothing accumulated before the first step
    ASSERT(integrator.getPotentialEnergyRefinement() == 0.0);
    const int n = 200;
    double sp = 0, sp2 = 0, sr = 0, sr2 = 0;
    for (int i = 0; i < n; i++) {
        integrator.step(1);
        State s = context.getState(State::Energy);
        vector<double> refined = integrator.getRefinedKineticEnergies();
        const double plain = s.getKineticEnergy() + s.getPotentialEnergy();
        const double shadow = refined[0] + refined[1] + s.getPotentialEnergy() + integrator.getPotentialEnergyRefinement();
        sp += plain; sp2 += plain*plain; sr += shadow; sr2 += shadow*shadow;
    }
    const double stdPlain = sqrt(sp2/n - (sp/n)*(sp/n)), stdRefined = sqrt(fabs(sr2/n - (sr/n)*(sr/n)));
    ASSERT(integrator.getPotentialEnergyRefinement() < 0.0);
    ASSERT(stdRefined < 0.3*stdPlain);
}

static void testErrors(Platform& platform) {
    System system;
    for (int i = 0; i < 4; i++) system.addParticle(12.0);
    {
        vector<int> wrong(3, 0);
        RigidBodyIntegrator integrator(0.001, wrong);
        bool thrown = false;
        try { Context context(system, integrator, platform); } catch (const OpenMMException& e) {
            thrown = string(e.what()).find("Number of body indices differs") != string::npos;
        }
        ASSERT(thrown);
    }
    vector<int> idx = {1, 1, 1, 0};
    RigidBodyIntegrator integrator(0.001, idx);
    bool thrown = false;
    try { integrator.setRotationMode(-2); } catch (const OpenMMException& e) { thrown = string(e.what()) == "Rotation mode cannot be negative"; }
    ASSERT(thrown);
    thrown = false;
    try { integrator.step(1); } catch (const OpenMMException& e) { thrown = string(e.what()).find("not bound to a context") != string::npos; }
    ASSERT(thrown);
    Context context(system, integrator, platform);
    thrown = false;
    try { integrator.setRotationMode(2); } catch (const OpenMMException& e) { thrown = string(e.what()).find("already bound to a context") != string::npos; }
    ASSERT(thrown);
    System constrained;
    for (int i = 0; i < 4; i++) constrained.addParticle(12.0);
    constrained.addConstraint(0, 3, 0.1);
    RigidBodyIntegrator integrator2(0.001, idx);
    thrown = false;
    try { Context c2(constrained, integrator2, platform); } catch (const OpenMMException& e) {
        thrown = string(e.what()) == "Constraints involving rigid-body atoms are not allowed";
    }
    ASSERT(thrown);
}
#endif
