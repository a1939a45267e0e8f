// Symmetric memory on the CUDA virtual-memory-management API with NVSwitch multicast objects.
//
// The substrate SURVEY §5.8 asks for: every rank cuMemCreate()s a physical allocation that can be exported as a POSIX
// file descriptor, peers import + map it (unicast peer pointers, like CUDA IPC but without the one-mapping-per-process
// limits), and one multicast object is bound to all the allocations so that a single `multimem.st` reaches every GPU and
// a single `multimem.ld_reduce` returns the in-switch sum over all GPUs (NVLS).  The file descriptors travel between the
// processes over unix sockets (python side: `parallel/symm_mem.py`).
//
// The driver API is resolved at run time through cudaGetDriverEntryPoint, so the library has no link-time dependency
// on libcuda (it must load on the GPU-less build box).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

namespace {

template <typename Fn> Fn drv(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<Fn>(fn);
}

#define RB_DRV(var, sym, type)              \
  static type var = drv<type>(sym);         \
  if (var == nullptr) return -100;

#define RB_CU(call, code)                                                          \
  do {                                                                             \
    CUresult _r = (call);                                                          \
    if (_r != CUDA_SUCCESS) {                                                      \
      fprintf(stderr, "[realhf_b200 vmm] %s failed: CUresult %d\n", #call, (int)_r); \
      return code;                                                                 \
    }                                                                              \
  } while (0)

typedef CUresult (*fnGetAttr)(int*, CUdevice_attribute, CUdevice);
typedef CUresult (*fnGran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
typedef CUresult (*fnCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult (*fnExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
typedef CUresult (*fnImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
typedef CUresult (*fnReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*fnMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*fnSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
typedef CUresult (*fnUnmap)(CUdeviceptr, size_t);
typedef CUresult (*fnAddrFree)(CUdeviceptr, size_t);
typedef CUresult (*fnRelease)(CUmemGenericAllocationHandle);
typedef CUresult (*fnMcCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
typedef CUresult (*fnMcAddDev)(CUmemGenericAllocationHandle, CUdevice);
typedef CUresult (*fnMcBind)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
typedef CUresult (*fnMcGran)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
typedef CUresult (*fnMcUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);

CUmemAllocationProp alloc_prop(int dev) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = dev;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

int map_rw(CUmemGenericAllocationHandle h, size_t size, size_t align, int dev, uint64_t* ptr_out) {
  RB_DRV(pReserve, "cuMemAddressReserve", fnReserve);
  RB_DRV(pMap, "cuMemMap", fnMap);
  RB_DRV(pAccess, "cuMemSetAccess", fnSetAccess);
  CUdeviceptr p = 0;
  RB_CU(pReserve(&p, size, align, 0, 0), -11);
  RB_CU(pMap(p, size, 0, h, 0), -12);
  CUmemAccessDesc a;
  memset(&a, 0, sizeof(a));
  a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  a.location.id = dev;
  a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  RB_CU(pAccess(p, size, &a, 1), -13);
  *ptr_out = (uint64_t)p;
  return 0;
}

}  // namespace

extern "C" {

// 1 if the device can bind memory to NVSwitch multicast objects (multimem.* instructions), 0 if not, <0 on error.
int rb_vmm_multicast_supported(int dev) {
  RB_DRV(pAttr, "cuDeviceGetAttribute", fnGetAttr);
  int v = 0;
  RB_CU(pAttr(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev), -1);
  return v;
}

// Allocation granularity that satisfies both cuMemCreate and (when `ndev` > 1) multicast binding for `ndev` devices.
int64_t rb_vmm_granularity(int dev, int ndev) {
  RB_DRV(pGran, "cuMemGetAllocationGranularity", fnGran);
  CUmemAllocationProp p = alloc_prop(dev);
  size_t g = 0;
  RB_CU(pGran(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), -1);
  if (ndev > 1 && rb_vmm_multicast_supported(dev) == 1) {
    static fnMcGran pMcGran = drv<fnMcGran>("cuMulticastGetGranularity");
    if (pMcGran != nullptr) {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof(mp));
      mp.numDevices = ndev;
      mp.size = g;
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t mg = 0;
      if (pMcGran(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
    }
  }
  return (int64_t)g;
}

// Physical allocation of `size` bytes (a multiple of the granularity) on `dev`, mapped read-write, exported as an fd.
int rb_vmm_alloc(int64_t size, int64_t align, int dev, uint64_t* ptr, uint64_t* handle, int* fd) {
  RB_DRV(pCreate, "cuMemCreate", fnCreate);
  RB_DRV(pExport, "cuMemExportToShareableHandle", fnExport);
  CUmemAllocationProp p = alloc_prop(dev);
  CUmemGenericAllocationHandle h;
  RB_CU(pCreate(&h, (size_t)size, &p, 0), -1);
  int f = -1;
  RB_CU(pExport(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), -2);
  int rc = map_rw(h, (size_t)size, (size_t)align, dev, ptr);
  if (rc != 0) return rc;
  *handle = (uint64_t)h;
  *fd = f;
  return 0;
}

// Map a peer's allocation (received as an fd) into this process for access from `dev`.
int rb_vmm_import(int fd, int64_t size, int64_t align, int dev, uint64_t* ptr, uint64_t* handle) {
  RB_DRV(pImport, "cuMemImportFromShareableHandle", fnImport);
  CUmemGenericAllocationHandle h;
  RB_CU(pImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), -1);
  int rc = map_rw(h, (size_t)size, (size_t)align, dev, ptr);
  if (rc != 0) return rc;
  *handle = (uint64_t)h;
  return 0;
}

int rb_vmm_free(uint64_t ptr, int64_t size, uint64_t handle) {
  RB_DRV(pUnmap, "cuMemUnmap", fnUnmap);
  RB_DRV(pFree, "cuMemAddressFree", fnAddrFree);
  RB_DRV(pRelease, "cuMemRelease", fnRelease);
  if (ptr) {
    RB_CU(pUnmap((CUdeviceptr)ptr, (size_t)size), -1);
    RB_CU(pFree((CUdeviceptr)ptr, (size_t)size), -2);
  }
  if (handle) RB_CU(pRelease((CUmemGenericAllocationHandle)handle), -3);
  return 0;
}

// Multicast object for `ndev` devices and `size` bytes; created by one rank, exported as an fd for the others.
int rb_mc_create(int64_t size, int ndev, uint64_t* mc, int* fd) {
  RB_DRV(pMcCreate, "cuMulticastCreate", fnMcCreate);
  RB_DRV(pExport, "cuMemExportToShareableHandle", fnExport);
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = ndev;
  mp.size = (size_t)size;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  RB_CU(pMcCreate(&h, &mp), -1);
  int f = -1;
  RB_CU(pExport(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), -2);
  *mc = (uint64_t)h;
  *fd = f;
  return 0;
}

int rb_mc_import(int fd, uint64_t* mc) {
  RB_DRV(pImport, "cuMemImportFromShareableHandle", fnImport);
  CUmemGenericAllocationHandle h;
  RB_CU(pImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), -1);
  *mc = (uint64_t)h;
  return 0;
}

// Every participating device must be added before any memory is bound.
int rb_mc_add_device(uint64_t mc, int dev) {
  RB_DRV(pAdd, "cuMulticastAddDevice", fnMcAddDev);
  RB_CU(pAdd((CUmemGenericAllocationHandle)mc, dev), -1);
  return 0;
}

// Bind this rank's physical allocation at offset 0 of the multicast object, then map the object (multicast pointer).
int rb_mc_bind_and_map(uint64_t mc, uint64_t mem_handle, int64_t size, int64_t align, int dev, uint64_t* mc_ptr) {
  RB_DRV(pBind, "cuMulticastBindMem", fnMcBind);
  RB_CU(pBind((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem_handle, 0, (size_t)size, 0), -1);
  return map_rw((CUmemGenericAllocationHandle)mc, (size_t)size, (size_t)align, dev, mc_ptr);
}

int rb_mc_unbind(uint64_t mc, int dev, int64_t size) {
  RB_DRV(pUnbind, "cuMulticastUnbind", fnMcUnbind);
  RB_CU(pUnbind((CUmemGenericAllocationHandle)mc, dev, 0, (size_t)size), -1);
  return 0;
}

int rb_close_fd(int fd) { return close(fd); }

}  // extern "C"
