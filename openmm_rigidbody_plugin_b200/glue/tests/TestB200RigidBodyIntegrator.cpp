// The glue's integration tests on the (shim) Reference platform: B200IntegrateRigidBodyStepKernel, host vectors in and out.
#include "RigidBodyTests.h"

int main() {
    try {
        ReferencePlatform* platform = new ReferencePlatform();
        Platform::registerPlatform(platform);
        registerRigidBodyB200KernelFactories();
        testErrors(*platform);
        testSingleBond(*platform);
        testRigidWaters(*platform, 0);
        testRigidWaters(*platform, 3);
        testConstrainedFreeAtoms(*platform);
        testRefinedEnergies(*platform);
    }
    catch (const exception& e) {
        cout << "exception: " << e.what() << endl;
        return 1;
    }
    cout << "Done" << endl;
    return 0;
}
