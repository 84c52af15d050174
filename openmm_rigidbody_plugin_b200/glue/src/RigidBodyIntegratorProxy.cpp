// XML proxy: version 1, stepSize, constraintTolerance, bodyIndices/bodyIndex@index - the node layout of
// serialization/src/RigidBodyIntegratorProxy.cpp:43-67 (rotation mode and refined-energy flag are not persisted there either).
#include "RigidBodyIntegratorProxy.h"
#include "RigidBodyIntegrator.h"
#include "openmm/OpenMMException.h"
#include <typeinfo>

using namespace RigidBodyPlugin;
using namespace OpenMM;

RigidBodyIntegratorProxy::RigidBodyIntegratorProxy() : SerializationProxy("RigidBodyIntegrator") {}

void RigidBodyIntegratorProxy::serialize(const void* object, SerializationNode& node) const {
    const RigidBodyIntegrator& integrator = *reinterpret_cast<const RigidBodyIntegrator*>(object);
    node.setIntProperty("version", 1);
    node.setDoubleProperty("stepSize", integrator.getStepSize());
    node.setDoubleProperty("constraintTolerance", integrator.getConstraintTolerance());
    SerializationNode& list = node.createChildNode("bodyIndices");
    const std::vector<int> indices = integrator.getBodyIndices();
    for (size_t i = 0; i < indices.size(); i++) list.createChildNode("bodyIndex").setIntProperty("index", indices[i]);
}

void* RigidBodyIntegratorProxy::deserialize(const SerializationNode& node) const {
    if (node.getIntProperty("version") != 1) throw OpenMMException("Unsupported version number");
    std::vector<int> indices;
    const std::vector<SerializationNode>& children = node.getChildNode("bodyIndices").getChildren();
    for (size_t i = 0; i < children.size(); i++) indices.push_back(children[i].getIntProperty("index"));
    RigidBodyIntegrator* integrator = new RigidBodyIntegrator(node.getDoubleProperty("stepSize"), indices);
    integrator->setConstraintTolerance(node.getDoubleProperty("constraintTolerance"));
    return integrator;
}

extern "C" void registerRigidBodySerializationProxies() {
    SerializationProxy::registerProxy(typeid(RigidBodyIntegrator), new RigidBodyIntegratorProxy());
}
