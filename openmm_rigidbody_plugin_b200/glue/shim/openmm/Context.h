#ifndef OPENMM_CONTEXT_H_
#define OPENMM_CONTEXT_H_
// shim, see Vec3.h: a Context bound to the (shim) Reference platform
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"
namespace OpenMM {
class Context {
public:
    Context(const System& system, Integrator& integrator, Platform& platform)
        : data(new ReferencePlatform::PlatformData(system)), impl(new ContextImpl(*this, system, integrator, &platform, data)) {
        integrator.initialize(*impl);
    }
    ~Context() { impl->getIntegrator().cleanup(); delete impl; delete data; }
    const System& getSystem() const { return impl->getSystem(); }
    Integrator& getIntegrator() { return impl->getIntegrator(); }
    Platform& getPlatform() { return impl->getPlatform(); }
    double getTime() const { return data->time; }
    void setPositions(const std::vector<Vec3>& positions) {
        if ((int) positions.size() != data->numParticles) throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
        *data->positions = positions;
        impl->getIntegrator().stateChanged(State::Positions);
    }
    void setVelocities(const std::vector<Vec3>& velocities) {
        if ((int) velocities.size() != data->numParticles) throw OpenMMException("Called setVelocities() on a Context with the wrong number of velocities");
        *data->velocities = velocities;
        impl->getIntegrator().stateChanged(State::Velocities);
    }
    State getState(int types) {
        State state;
        state.time = data->time;
        if (types & State::Energy) {
            state.pe = impl->calcForcesAndEnergy(false, true);
            state.ke = impl->getIntegrator().computeKineticEnergy();
        }
        if (types & State::Forces) { impl->calcForcesAndEnergy(true, false); state.forces = *data->forces; }
        if (types & State::Positions) state.positions = *data->positions;
        if (types & State::Velocities) state.velocities = *data->velocities;
        return state;
    }
private:
    ReferencePlatform::PlatformData* data;
    ContextImpl* impl;
};

}
#endif
