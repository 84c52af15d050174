#ifndef OPENMM_SERIALIZATIONPROXY_H_
#define OPENMM_SERIALIZATIONPROXY_H_
// shim, see Vec3.h
#include "SerializationNode.h"
#include <map>
#include <string>
#include <typeinfo>
namespace OpenMM {
class SerializationProxy {
public:
    SerializationProxy(const std::string& typeName) : typeName(typeName) {}
    virtual ~SerializationProxy() {}
    const std::string& getTypeName() const { return typeName; }
    virtual void serialize(const void* object, SerializationNode& node) const = 0;
    virtual void* deserialize(const SerializationNode& node) const = 0;
    static void registerProxy(const std::type_info& type, const SerializationProxy* proxy) {
        byType()[type.name()] = proxy;
        byName()[proxy->getTypeName()] = proxy;
    }
    static const SerializationProxy& getProxy(const std::string& typeName) {
        std::map<std::string, const SerializationProxy*>::const_iterator it = byName().find(typeName);
        if (it == byName().end()) throw OpenMMException("There is no serialization proxy registered for type " + typeName);
        return *it->second;
    }
    static const SerializationProxy& getProxy(const std::type_info& type) {
        std::map<std::string, const SerializationProxy*>::const_iterator it = byType().find(type.name());
        if (it == byType().end()) throw OpenMMException(std::string("There is no serialization proxy registered for type ") + type.name());
        return *it->second;
    }
private:
    static std::map<std::string, const SerializationProxy*>& byType() { static std::map<std::string, const SerializationProxy*> m; return m; }
    static std::map<std::string, const SerializationProxy*>& byName() { static std::map<std::string, const SerializationProxy*> m; return m; }
    std::string typeName;
};
}
#endif
