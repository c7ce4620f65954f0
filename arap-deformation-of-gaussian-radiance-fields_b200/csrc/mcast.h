// NVSwitch multicast memory (csrc/mcast.cpp): host-side set-up of one multicast object over the ranks of a node.
#pragma once
#include <cstddef>

namespace arapgs {

constexpr int ARAP_MCAST_NAME = 64;
struct Mcast {
  unsigned long long mc = 0, mem = 0;   // CUmemGenericAllocationHandle: the multicast object, this rank's bound allocation
  void* local = nullptr;                // this rank's memory (ordinary loads / stores)
  void* mc_ptr = nullptr;               // the multicast mapping: multimem.st here lands in every rank's `local`
  size_t size = 0;
  int export_fd = -1, listen_fd = -1, device = 0;
  bool local_mapped = false, mc_mapped = false, bound = false;
};
int mcast_supported(int device);                                                             // CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
int mcast_size(size_t bytes, int world, int device, size_t* rounded);                        // bytes rounded up to the granularities
int mcast_root_begin(Mcast* m, size_t size, int world, char name_out[ARAP_MCAST_NAME]);      // rank 0: create, export, listen
int mcast_root_serve(Mcast* m, int n_peers);                                                 // rank 0: hand the descriptor to n peers
int mcast_peer_join(Mcast* m, size_t size, const char* name);                                // other ranks: receive + import
int mcast_bind_and_map(Mcast* m, int device);                                                // all ranks
void mcast_destroy(Mcast* m);

}  // namespace arapgs
