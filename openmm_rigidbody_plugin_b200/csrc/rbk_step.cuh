// rbk_step.cuh - the per-body and per-atom pieces of the integrator step, as host/device inline
// functions.  The CUDA kernels (rbk_kernels.cu) wire these into tiles; the CPU unit tests call the
// very same functions on the host to check the arithmetic against the oracle.
//
//   bodyPart1 : RigidBodySystem::integratePart1 body loop      openmmapi/src/RigidBodySystem.cpp:177-186
//   atomPosition : RigidBody::updateAtomicPositions             openmmapi/src/RigidBody.cpp:148-153
//   atomForceTorque + bodyPart2 : forceAndTorque + second kick + updateAtomicVelocities
//                                                               RigidBody.cpp:159-183, RigidBodySystem.cpp:198-203
//   freePart1 / freePart2 : free-atom velocity Verlet           RigidBodySystem.cpp:172-176, 196-197
//   bodyKinetic / freeKinetic : computeKineticEnergies          RigidBodySystem.cpp:210-220
#pragma once
#include "rbk_math.cuh"

namespace rbk {

// First half kick, drift and free rotation of one body.  tau is the space-frame torque; the
// quaternion-frame torque the reference stores is C(q) tau with the SAME q (q does not change
// between Part 2 and the next Part 1), so it is rebuilt here instead of being stored.
template <bool EXACT>
RBK_HD void bodyPart1(double dt, int nSplit, d3 F, d3 tau, double invm, d3 invI, d3& r, d3& p, d4& q, d4& pi) {
    const double halfDt = 0.5*dt;
    p = p + F*halfDt;
    pi = pi + quatC(q, tau)*dt;
    r = r + p*(invm*dt);
    if (EXACT) exactRotation(dt, invI, q, pi);
    else noSquish(dt, nSplit, invI, q, pi);
}

// Part 1 with the exact rotation's series order taken from the ladder (rbk_math.cuh): rung = 0, 1, 2
RBK_HD void bodyPart1Ladder(int rung, double dt, d3 F, d3 tau, double invm, d3 invI, d3& r, d3& p, d4& q, d4& pi, unsigned& flags) {
    p = p + F*(0.5*dt);
    pi = pi + quatC(q, tau)*dt;
    r = r + p*(invm*dt);
    exactRotationLadder(rung, dt, invI, q, pi, flags);
}

RBK_HD d3 atomPosition(d3 r, d4 q, d3 d) { return r + bodyToSpace(q, d); }

// Second half kick of one body from the reduced force / torque; returns what the atoms need to
// rebuild their velocities: v_cm and the space-frame angular velocity.
RBK_HD void bodyPart2(double dt, d3 F, d3 tau, double invm, d3 invI, d4 q, d3& p, d4& pi, d3& vcm, d3& omegaSpace) {
    const double halfDt = 0.5*dt;
    p = p + F*halfDt;
    pi = pi + quatC(q, tau)*dt;
    const d3 L = quatBt(q, pi)*0.5;
    const d3 omega = {invI.x*L.x, invI.y*L.y, invI.z*L.z};
    omegaSpace = bodyToSpace(q, omega);
    vcm = p*invm;
}

RBK_HD d3 atomVelocity(d3 vcm, d3 omegaSpace, d3 delta) { return vcm + cross(omegaSpace, delta); }

RBK_HD void freePart1(double dt, d3 f, double invm, d3& x, d3& v) {
    v = v + f*invm*(0.5*dt);
    x = x + v*dt;
}

RBK_HD void freePart2(double dt, d3 f, double invm, d3 x, d3 saved, d3& v) {
    v = v + (f*invm*(0.5*dt) + (x - saved)*(1.0/dt));
}

RBK_HD void bodyKinetic(d3 p, d4 q, d4 pi, double invm, d3 invI, double& twoKt, double& twoKr) {
    const d3 vcm = p*invm;
    const d3 L = quatBt(q, pi)*0.5;
    const d3 omega = {invI.x*L.x, invI.y*L.y, invI.z*L.z};
    twoKt = dot(p, vcm);
    twoKr = dot(L, omega);
}

RBK_HD double freeKinetic(d3 v, double invm) { return dot(v, v)/invm; }

} // namespace rbk
