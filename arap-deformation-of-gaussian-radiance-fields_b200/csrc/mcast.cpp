// NVSwitch multicast ("NVLS") memory for the fused multi-GPU exchange (arap_comm_set_mode(2)).
//
// One multicast object spans the ranks of a node; every rank binds a local allocation of the same size to it, and a
// `multimem.st` to the object's address is replicated by the switch into all of them — the apply kernel's epilogue then sends
// each deformed Gaussian's pose ONCE (40 bytes of NVLink egress) instead of once per peer.  Host side only, CUDA driver VMM
// API through cudaGetDriverEntryPoint (libarapgs has no link-time dependency on libcuda):
//   rank 0:   cuMulticastCreate -> POSIX file descriptor -> handed to the other processes over a unix socket (SCM_RIGHTS)
//   everyone: cuMemImportFromShareableHandle, cuMulticastAddDevice, cuMemCreate (shareable, as bound memory must be),
//             map it locally, cuMulticastBindMem, map the multicast object.
// The driver blocks binds and mappings until all devices have been added, so no extra barrier is needed in between.
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include "mcast.h"
#include "../../include/arapgs.h"

namespace arapgs {
void set_error(const std::string& msg);

namespace {
struct Drv {
  bool ok = false;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
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
};
Drv g_drv;

template <typename F>
bool load(const char* name, F& f) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !p) return false;
  f = reinterpret_cast<F>(p);
  return true;
}
bool drv_load() {
  if (g_drv.ok) return true;
  Drv d;
  bool ok = load("cuGetErrorString", d.GetErrorString) && load("cuDeviceGetAttribute", d.DeviceGetAttribute) &&
            load("cuMulticastCreate", d.MulticastCreate) && load("cuMulticastAddDevice", d.MulticastAddDevice) &&
            load("cuMulticastBindMem", d.MulticastBindMem) && load("cuMulticastUnbind", d.MulticastUnbind) &&
            load("cuMulticastGetGranularity", d.MulticastGetGranularity) && load("cuMemCreate", d.MemCreate) &&
            load("cuMemRelease", d.MemRelease) && load("cuMemAddressReserve", d.MemAddressReserve) &&
            load("cuMemAddressFree", d.MemAddressFree) && load("cuMemMap", d.MemMap) && load("cuMemUnmap", d.MemUnmap) &&
            load("cuMemSetAccess", d.MemSetAccess) && load("cuMemGetAllocationGranularity", d.MemGetAllocationGranularity) &&
            load("cuMemExportToShareableHandle", d.MemExportToShareableHandle) &&
            load("cuMemImportFromShareableHandle", d.MemImportFromShareableHandle);
  if (!ok) { cudaGetLastError(); return false; }
  d.ok = true; g_drv = d;
  return true;
}
int fail(const char* what, CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  set_error(std::string("multicast: ") + what + ": " + (s ? s : "driver error") + " (" + std::to_string((int)r) + ")");
  return ARAP_ERR_UNSUPPORTED;
}
int fail_errno(const char* what) { set_error(std::string("multicast: ") + what + ": " + strerror(errno)); return ARAP_ERR_UNSUPPORTED; }
#define DRV_TRY(call, what) do { CUresult _r = (call); if (_r != CUDA_SUCCESS) return fail(what, _r); } while (0)

CUmulticastObjectProp mc_prop(int world, size_t size) {
  CUmulticastObjectProp p; memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned)world; p.size = size; p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR; p.flags = 0;
  return p;
}
CUmemAllocationProp mem_prop(int device) {
  CUmemAllocationProp p; memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED; p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE; p.location.id = device;
  return p;
}
socklen_t sock_addr(const char* name, sockaddr_un& a) {   // abstract namespace: no file to clean up
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  const size_t n = strlen(name);
  memcpy(a.sun_path + 1, name, n);
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}
}  // namespace

int mcast_supported(int device) {
  if (!drv_load()) return 0;
  int v = 0;
  if (g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, (CUdevice)device) != CUDA_SUCCESS) return 0;
  return v;
}

int mcast_size(size_t bytes, int world, int device, size_t* out) {
  if (!drv_load()) { set_error("multicast: driver entry points unavailable"); return ARAP_ERR_UNSUPPORTED; }
  size_t g1 = 0, g2 = 0;
  CUmulticastObjectProp mp = mc_prop(world, bytes);
  DRV_TRY(g_drv.MulticastGetGranularity(&g1, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
  CUmemAllocationProp ap = mem_prop(device);
  DRV_TRY(g_drv.MemGetAllocationGranularity(&g2, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  const size_t g = g1 > g2 ? g1 : g2;
  *out = (bytes + g - 1) / g * g;
  return ARAP_OK;
}

int mcast_root_begin(Mcast* m, size_t size, int world, char name_out[ARAP_MCAST_NAME]) {
  if (!drv_load()) { set_error("multicast: driver entry points unavailable"); return ARAP_ERR_UNSUPPORTED; }
  CUmulticastObjectProp mp = mc_prop(world, size);
  CUmemGenericAllocationHandle h = 0;
  DRV_TRY(g_drv.MulticastCreate(&h, &mp), "cuMulticastCreate");
  m->mc = (unsigned long long)h; m->size = size;
  int fd = -1;
  DRV_TRY(g_drv.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
  m->export_fd = fd;
  static int counter = 0;
  snprintf(name_out, ARAP_MCAST_NAME, "arapgs-mc-%d-%d", (int)getpid(), counter++);
  const int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) return fail_errno("socket");
  sockaddr_un a; const socklen_t len = sock_addr(name_out, a);
  if (bind(s, (sockaddr*)&a, len) != 0 || listen(s, 16) != 0) { close(s); return fail_errno("bind/listen"); }
  timeval tv; tv.tv_sec = 60; tv.tv_usec = 0;
  setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));   // accept() gives up if a peer never shows up
  m->listen_fd = s;
  return ARAP_OK;
}

int mcast_root_serve(Mcast* m, int n_peers) {
  for (int i = 0; i < n_peers; i++) {
    const int c = accept(m->listen_fd, nullptr, nullptr);
    if (c < 0) return fail_errno("accept");
    char byte = 'm';
    iovec iov{&byte, 1};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    msghdr msg; memset(&msg, 0, sizeof(msg));
    msg.msg_iov = &iov; msg.msg_iovlen = 1; msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
    cmsghdr* cm = CMSG_FIRSTHDR(&msg);
    cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(cm), &m->export_fd, sizeof(int));
    const ssize_t r = sendmsg(c, &msg, 0);
    close(c);
    if (r != 1) return fail_errno("sendmsg");
  }
  close(m->listen_fd); m->listen_fd = -1;
  return ARAP_OK;
}

int mcast_peer_join(Mcast* m, size_t size, const char* name) {
  if (!drv_load()) { set_error("multicast: driver entry points unavailable"); return ARAP_ERR_UNSUPPORTED; }
  const int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) return fail_errno("socket");
  sockaddr_un a; const socklen_t len = sock_addr(name, a);
  int tries = 0;
  while (connect(s, (sockaddr*)&a, len) != 0) {   // the root listens before the name is published; retry only on a transient refusal
    if (++tries > 200) { close(s); return fail_errno("connect"); }
    usleep(10000);
  }
  char byte = 0;
  iovec iov{&byte, 1};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg; memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov; msg.msg_iovlen = 1; msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  const ssize_t r = recvmsg(s, &msg, 0);
  close(s);
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  if (r != 1 || !cm || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) { set_error("multicast: no file descriptor received from rank 0"); return ARAP_ERR_UNSUPPORTED; }
  int fd = -1;
  memcpy(&fd, CMSG_DATA(cm), sizeof(int));
  CUmemGenericAllocationHandle h = 0;
  const CUresult ir = g_drv.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  close(fd);
  if (ir != CUDA_SUCCESS) return fail("cuMemImportFromShareableHandle", ir);
  m->mc = (unsigned long long)h; m->size = size;
  return ARAP_OK;
}

int mcast_bind_and_map(Mcast* m, int device) {
  const CUmemGenericAllocationHandle mc = (CUmemGenericAllocationHandle)m->mc;
  DRV_TRY(g_drv.MulticastAddDevice(mc, (CUdevice)device), "cuMulticastAddDevice");
  CUmemAllocationProp ap = mem_prop(device);
  CUmemGenericAllocationHandle mem = 0;
  DRV_TRY(g_drv.MemCreate(&mem, m->size, &ap, 0), "cuMemCreate");
  m->mem = (unsigned long long)mem;
  CUmemAccessDesc acc; memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = device; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CUdeviceptr va = 0;
  DRV_TRY(g_drv.MemAddressReserve(&va, m->size, 0, 0, 0), "cuMemAddressReserve");
  m->local = (void*)va;
  DRV_TRY(g_drv.MemMap(va, m->size, 0, mem, 0), "cuMemMap(local)");
  m->local_mapped = true;
  DRV_TRY(g_drv.MemSetAccess(va, m->size, &acc, 1), "cuMemSetAccess(local)");
  DRV_TRY(g_drv.MulticastBindMem(mc, 0, mem, 0, m->size, 0), "cuMulticastBindMem");   // blocks until every device has been added
  m->bound = true; m->device = device;
  CUdeviceptr mva = 0;
  DRV_TRY(g_drv.MemAddressReserve(&mva, m->size, 0, 0, 0), "cuMemAddressReserve(mc)");
  m->mc_ptr = (void*)mva;
  DRV_TRY(g_drv.MemMap(mva, m->size, 0, mc, 0), "cuMemMap(mc)");
  m->mc_mapped = true;
  DRV_TRY(g_drv.MemSetAccess(mva, m->size, &acc, 1), "cuMemSetAccess(mc)");
  return ARAP_OK;
}

void mcast_destroy(Mcast* m) {
  if (!g_drv.ok) return;
  if (m->mc_ptr) { if (m->mc_mapped) g_drv.MemUnmap((CUdeviceptr)m->mc_ptr, m->size); g_drv.MemAddressFree((CUdeviceptr)m->mc_ptr, m->size); }
  if (m->bound) g_drv.MulticastUnbind((CUmemGenericAllocationHandle)m->mc, (CUdevice)m->device, 0, m->size);
  if (m->local) { if (m->local_mapped) g_drv.MemUnmap((CUdeviceptr)m->local, m->size); g_drv.MemAddressFree((CUdeviceptr)m->local, m->size); }
  if (m->mem) g_drv.MemRelease((CUmemGenericAllocationHandle)m->mem);
  if (m->mc) g_drv.MemRelease((CUmemGenericAllocationHandle)m->mc);
  if (m->export_fd >= 0) close(m->export_fd);
  if (m->listen_fd >= 0) close(m->listen_fd);
  *m = Mcast{};
}

}  // namespace arapgs
