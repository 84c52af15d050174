// XML round trip of the integrator, as serialization/tests/TestSerializeRigidBodyIntegrator.cpp:48-63 of the reference.
#include "RigidBodyIntegrator.h"
#include "RigidBodyIntegratorProxy.h"
#include "openmm/serialization/XmlSerializer.h"
#include <iostream>
#include <sstream>
#include <stdexcept>

using namespace RigidBodyPlugin;
using namespace OpenMM;
using namespace std;

int main() {
    try {
        registerRigidBodySerializationProxies();
        const int indices[] = {5, 4, 3, 2, 1};
        vector<int> bodyIndices(indices, indices + 5);
        RigidBodyIntegrator original(0.00342, bodyIndices);
        original.setConstraintTolerance(3e-6);
        stringstream buffer;
        XmlSerializer::serialize<Integrator>(&original, "RigidBodyIntegrator", buffer);
        RigidBodyIntegrator* copy = dynamic_cast<RigidBodyIntegrator*>(XmlSerializer::deserialize<Integrator>(buffer));
        if (copy == NULL) throw runtime_error("deserialized object is not a RigidBodyIntegrator");
        if (copy->getConstraintTolerance() != original.getConstraintTolerance()) throw runtime_error("constraintTolerance differs");
        if (copy->getStepSize() != original.getStepSize()) throw runtime_error("stepSize differs");
        if (copy->getBodyIndices() != bodyIndices) throw runtime_error("bodyIndices differ");
        delete copy;
        string xml = buffer.str();
        if (xml.find("version=\"1\"") == string::npos || xml.find("<bodyIndices>") == string::npos) throw runtime_error("unexpected XML layout");
    }
    catch (const exception& e) {
        cout << "exception: " << e.what() << endl;
        return 1;
    }
    cout << "Done" << endl;
    return 0;
}
