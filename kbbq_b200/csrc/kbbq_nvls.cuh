// kbbq_nvls.cuh -- the multi-GPU filter merge through NVSwitch multicast (NVLS): in-switch bitwise OR.
//
// or_allreduce_p2p_kernel moves 2(N-1)/N of the bit array per GPU and direction over NVLink (every GPU reads its slice
// from N-1 peers and writes the result to N-1 peers): 29 ms of a 243 ms step at N = 8 on the config-3 shards, at the
// link's bandwidth (685 GB/s per direction).  With the N copies of the array bound to one multicast object the switch
// does both halves: `multimem.ld_reduce.or` fetches a word from all N copies and ORs them in the switch (1/N of the
// array arrives per GPU), `multimem.st` writes the result to all N copies (1/N leaves).
//
// Needs: filters allocated with the virtual-memory API as shareable (POSIX fd) memory (KBBQ_FILTER_VMM=1 at
// kbbq_bloom_create), a multicast object created by one rank and imported by the others (the fd travels over a Unix
// socket: kbbq_b200/dist.py), every device added, then every rank's memory bound and the object mapped.  The driver
// API is reached through cudaGetDriverEntryPoint: the library does not link libcuda and still loads on a box without a
// driver.  Everything here fails soft: any error makes the caller fall back to another merge.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

namespace kbbq {
namespace nvls {

struct Driver {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    bool ok = false;
    std::string why;
};

inline const Driver& driver() {
    static Driver d = [] {
        Driver x;
        auto get = [&](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) {
                cudaGetLastError();
                if (x.why.empty()) x.why = std::string("driver entry point ") + name + " not available";
                return false;
            }
            return true;
        };
        bool all = true;
#define KBBQ_GET(sym, member) all = get(sym, (void**)&x.member) && all
        KBBQ_GET("cuMemCreate", MemCreate);
        KBBQ_GET("cuMemRelease", MemRelease);
        KBBQ_GET("cuMemAddressReserve", MemAddressReserve);
        KBBQ_GET("cuMemAddressFree", MemAddressFree);
        KBBQ_GET("cuMemMap", MemMap);
        KBBQ_GET("cuMemUnmap", MemUnmap);
        KBBQ_GET("cuMemSetAccess", MemSetAccess);
        KBBQ_GET("cuMemGetAllocationGranularity", MemGetAllocationGranularity);
        KBBQ_GET("cuMemExportToShareableHandle", MemExportToShareableHandle);
        KBBQ_GET("cuMemImportFromShareableHandle", MemImportFromShareableHandle);
        KBBQ_GET("cuMulticastCreate", MulticastCreate);
        KBBQ_GET("cuMulticastAddDevice", MulticastAddDevice);
        KBBQ_GET("cuMulticastBindMem", MulticastBindMem);
        KBBQ_GET("cuMulticastUnbind", MulticastUnbind);
        KBBQ_GET("cuMulticastGetGranularity", MulticastGetGranularity);
        KBBQ_GET("cuDeviceGet", DeviceGet);
        KBBQ_GET("cuDeviceGetAttribute", DeviceGetAttribute);
        KBBQ_GET("cuGetErrorString", GetErrorString);
#undef KBBQ_GET
        x.ok = all;
        return x;
    }();
    return d;
}

inline std::string cu_error(const char* what, CUresult r) {
    const char* s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &s);
    return std::string(what) + ": " + (s ? s : "unknown driver error") + " (" + std::to_string((int)r) + ")";
}

inline CUmemAllocationProp alloc_prop(int device) {
    CUmemAllocationProp p = {};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    return p;
}

// the size a filter of `bytes` occupies as VMM memory: a multiple of the allocation granularity and of the multicast
// granularity (2 MiB-ish each), so that the whole allocation can be bound to a multicast object later
inline bool padded_size(int device, size_t bytes, size_t& padded, std::string& why) {
    const Driver& d = driver();
    if (!d.ok) { why = d.why; return false; }
    const CUmemAllocationProp p = alloc_prop(device);
    size_t g = 0, mg = 0;
    CUresult r = d.MemGetAllocationGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (r != CUDA_SUCCESS) { why = cu_error("cuMemGetAllocationGranularity", r); return false; }
    CUmulticastObjectProp mp = {};
    mp.numDevices = 2;
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (d.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || mg == 0) mg = g;
    size_t step = g > mg ? g : mg;
    if (step % g || step % mg) step = g * mg;  // both are powers of two in practice
    padded = (bytes + step - 1) / step * step;
    return true;
}

// rank r ORs slice r of the array across all copies in the switch and writes it back to all copies
__global__ void __launch_bounds__(256) or_allreduce_mc_kernel(unsigned long long* mc, int rank, int world, unsigned long long n_words) {
    const unsigned long long per = (n_words + world - 1) / world;
    const unsigned long long lo = per * rank, hi = lo + per < n_words ? lo + per : n_words;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        unsigned long long v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.or.b64 %0, [%1];" : "=l"(v) : "l"(mc + i) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.b64 [%0], %1;" ::"l"(mc + i), "l"(v) : "memory");
    }
    __threadfence_system();
}

}  // namespace nvls
}  // namespace kbbq
