#ifndef RBK_GLUE_RIGIDBODYINTEGRATORPROXY_H_
#define RBK_GLUE_RIGIDBODYINTEGRATORPROXY_H_
// XML proxy of the integrator (serialization/include/RigidBodyIntegratorProxy.h of the reference).
#include "openmm/serialization/SerializationProxy.h"

namespace RigidBodyPlugin {

class RigidBodyIntegratorProxy : public OpenMM::SerializationProxy {
public:
    RigidBodyIntegratorProxy();
    void serialize(const void* object, OpenMM::SerializationNode& node) const;
    void* deserialize(const OpenMM::SerializationNode& node) const;
};

} // namespace RigidBodyPlugin

extern "C" void registerRigidBodySerializationProxies();
#endif
