#ifndef OPENMM_PLATFORM_H_
#define OPENMM_PLATFORM_H_
// shim, see Vec3.h
#include "Kernel.h"
#include "KernelFactory.h"
#include "OpenMMException.h"
#include <map>
#include <string>
#include <vector>
#ifndef OPENMM_EXPORT
#define OPENMM_EXPORT
#endif
namespace OpenMM {
class Platform {
public:
    virtual ~Platform() {
        for (std::map<std::string, KernelFactory*>::iterator it = factories.begin(); it != factories.end(); ++it) {
            bool shared = false;            // one factory may serve several kernel names
            for (std::map<std::string, KernelFactory*>::iterator jt = factories.begin(); jt != it; ++jt)
                if (jt->second == it->second) shared = true;
            if (!shared) delete it->second;
        }
    }
    virtual const std::string& getName() const = 0;
    virtual void contextCreated(ContextImpl& context) const {}
    virtual void contextDestroyed(ContextImpl& context) const {}
    void registerKernelFactory(const std::string& name, KernelFactory* factory) { factories[name] = factory; }
    void setPropertyDefaultValue(const std::string& property, const std::string& value) { properties[property] = value; }
    const std::string& getPropertyDefaultValue(const std::string& property) const {
        std::map<std::string, std::string>::const_iterator it = properties.find(property);
        if (it == properties.end()) throw OpenMMException("getPropertyDefaultValue: Illegal property name");
        return it->second;
    }
    bool supportsKernels(const std::vector<std::string>& kernelNames) const {
        for (size_t i = 0; i < kernelNames.size(); i++) if (factories.find(kernelNames[i]) == factories.end()) return false;
        return true;
    }
    Kernel createKernel(const std::string& name, ContextImpl& context) const {
        std::map<std::string, KernelFactory*>::const_iterator it = factories.find(name);
        if (it == factories.end()) throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
        return Kernel(it->second->createKernelImpl(name, *this, context));
    }
    static void registerPlatform(Platform* platform) { getPlatforms().push_back(platform); }
    static int getNumPlatforms() { return (int) getPlatforms().size(); }
    static Platform& getPlatform(int index) { return *getPlatforms()[index]; }
    static Platform& getPlatformByName(const std::string& name) {
        for (size_t i = 0; i < getPlatforms().size(); i++) if (getPlatforms()[i]->getName() == name) return *getPlatforms()[i];
        throw OpenMMException("There is no registered Platform called \"" + name + "\"");
    }
private:
    static std::vector<Platform*>& getPlatforms() { static std::vector<Platform*> platforms; return platforms; }
    std::map<std::string, KernelFactory*> factories;
    std::map<std::string, std::string> properties;
};
}
#endif
