#ifndef OPENMM_CONTEXT_H_
#define OPENMM_CONTEXT_H_
// shim, see Vec3.h: a Context bound to the (shim) Reference platform or to the (shim) CUDA platform
#include "openmm/internal/ContextImpl.h"
#include "openmm/cuda/CudaPlatform.h"
#include "openmm/reference/ReferencePlatform.h"
namespace OpenMM {
class Context {
public:
    Context(const System& system, Integrator& integrator, Platform& platform) : data(NULL), cudaData(NULL) {
        CudaPlatform* cuda = dynamic_cast<CudaPlatform*>(&platform);
        if (cuda != NULL) {
            cudaData = new CudaPlatform::PlatformData(system, platform.getPropertyDefaultValue(CudaPlatform::CudaPrecision()));
            impl = new ContextImpl(*this, system, integrator, &platform, cudaData, true);
        }
        else {
            data = new ReferencePlatform::PlatformData(system);
            impl = new ContextImpl(*this, system, integrator, &platform, data, false);
        }
        numParticles = system.getNumParticles();
        integrator.initialize(*impl);
    }
    ~Context() { impl->getIntegrator().cleanup(); delete impl; delete data; delete cudaData; }
    const System& getSystem() const { return impl->getSystem(); }
    Integrator& getIntegrator() { return impl->getIntegrator(); }
    Platform& getPlatform() { return impl->getPlatform(); }
    double getTime() const { return cudaData ? cudaData->contexts[0]->getTime() : data->time; }
    void setPositions(const std::vector<Vec3>& positions) {
        if ((int) positions.size() != numParticles) throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
        if (cudaData) cudaData->contexts[0]->uploadPositions(positions); else *data->positions = positions;
        impl->getIntegrator().stateChanged(State::Positions);
    }
    void setVelocities(const std::vector<Vec3>& velocities) {
        if ((int) velocities.size() != numParticles) throw OpenMMException("Called setVelocities() on a Context with the wrong number of velocities");
        if (cudaData) cudaData->contexts[0]->uploadVelocities(velocities); else *data->velocities = velocities;
        impl->getIntegrator().stateChanged(State::Velocities);
    }
    State getState(int types) {
        State state;
        state.time = getTime();
        if (types & State::Energy) {
            state.pe = impl->calcForcesAndEnergy(false, true);
            state.ke = impl->getIntegrator().computeKineticEnergy();
        }
        if (types & State::Forces) { impl->calcForcesAndEnergy(true, false); impl->getForces(state.forces); }
        if (types & State::Positions) impl->getPositions(state.positions);
        if (types & State::Velocities) impl->getVelocities(state.velocities);
        return state;
    }
private:
    ReferencePlatform::PlatformData* data;
    CudaPlatform::PlatformData* cudaData;
    ContextImpl* impl;
    int numParticles;
};

}
#endif
