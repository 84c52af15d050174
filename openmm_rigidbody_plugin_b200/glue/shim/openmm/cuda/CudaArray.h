#ifndef OPENMM_CUDAARRAY_H_
#define OPENMM_CUDAARRAY_H_
// shim, see ../Vec3.h: a device array with the interface of OpenMM's CudaArray that the rigid-body plugin uses
// (upload, download, getDevicePointer, getSize), backed by the CUDA runtime.
#include "openmm/OpenMMException.h"
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <string>
#include <vector>
namespace OpenMM {
class CudaContext;
class CudaArray {
public:
    CudaArray(CudaContext&, int size, int elementSize, const std::string& name) : pointer(0), size(size), elementSize(elementSize), name(name) {
        void* p = NULL;
        if (cudaMalloc(&p, (size_t) (size > 0 ? size : 1)*elementSize) != cudaSuccess) throw OpenMMException("Error creating array " + name);
        cudaMemset(p, 0, (size_t) (size > 0 ? size : 1)*elementSize);
        pointer = (CUdeviceptr) p;
    }
    ~CudaArray() { cudaFree((void*) pointer); }
    int getSize() const { return size; }
    int getElementSize() const { return elementSize; }
    const std::string& getName() const { return name; }
    CUdeviceptr& getDevicePointer() { return pointer; }
    void upload(const void* data, bool blocking = true) {
        if (cudaMemcpy((void*) pointer, data, (size_t) size*elementSize, cudaMemcpyHostToDevice) != cudaSuccess) throw OpenMMException("Error uploading array " + name);
    }
    void download(void* data, bool blocking = true) const {
        if (cudaMemcpy(data, (const void*) pointer, (size_t) size*elementSize, cudaMemcpyDeviceToHost) != cudaSuccess) throw OpenMMException("Error downloading array " + name);
    }
    template <class T> void upload(const std::vector<T>& data) { upload((const void*) &data[0]); }
    template <class T> void download(std::vector<T>& data) const { data.resize(size); download((void*) &data[0]); }
private:
    CudaArray(const CudaArray&);
    CudaArray& operator=(const CudaArray&);
    CUdeviceptr pointer;
    int size, elementSize;
    std::string name;
};
}
#endif
