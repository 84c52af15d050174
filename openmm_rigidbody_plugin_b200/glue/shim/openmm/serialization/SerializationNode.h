#ifndef OPENMM_SERIALIZATIONNODE_H_
#define OPENMM_SERIALIZATIONNODE_H_
// shim, see Vec3.h: property tree with the subset of OpenMM::SerializationNode the integrator proxy uses
#include "openmm/OpenMMException.h"
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class SerializationNode {
public:
    const std::string& getName() const { return name; }
    void setName(const std::string& n) { name = n; }
    const std::vector<SerializationNode>& getChildren() const { return children; }
    std::vector<SerializationNode>& getChildren() { return children; }
    const SerializationNode& getChildNode(const std::string& n) const {
        for (size_t i = 0; i < children.size(); i++) if (children[i].name == n) return children[i];
        throw OpenMMException("Unknown child '" + n + "' in node '" + name + "'");
    }
    SerializationNode& createChildNode(const std::string& n) { children.push_back(SerializationNode()); children.back().name = n; return children.back(); }
    const std::map<std::string, std::string>& getProperties() const { return properties; }
    bool hasProperty(const std::string& n) const { return properties.find(n) != properties.end(); }
    const std::string& getStringProperty(const std::string& n) const {
        std::map<std::string, std::string>::const_iterator it = properties.find(n);
        if (it == properties.end()) throw OpenMMException("Unknown property '" + n + "' in node '" + name + "'");
        return it->second;
    }
    SerializationNode& setStringProperty(const std::string& n, const std::string& v) { properties[n] = v; return *this; }
    int getIntProperty(const std::string& n) const { return std::atoi(getStringProperty(n).c_str()); }
    SerializationNode& setIntProperty(const std::string& n, int v) { char b[32]; std::snprintf(b, sizeof b, "%d", v); properties[n] = b; return *this; }
    double getDoubleProperty(const std::string& n) const { return std::strtod(getStringProperty(n).c_str(), NULL); }
    SerializationNode& setDoubleProperty(const std::string& n, double v) { char b[64]; std::snprintf(b, sizeof b, "%.17g", v); properties[n] = b; return *this; }
private:
    std::string name;
    std::vector<SerializationNode> children;
    std::map<std::string, std::string> properties;
};
}
#endif
