// emvs_internal.h — declarations shared by the translation units of libemvs_b200.so.
#ifndef EMVS_INTERNAL_H_
#define EMVS_INTERNAL_H_

#include "emvs_b200.h"

#include <cstddef>
#include <cstdint>

namespace emvs {

// thread-local error message behind emvs_last_error()
void set_error(const char* fmt, ...);

// host_geometry.cpp
void host_depth_vector(const emvs_shape& shape, float* out);
void host_virtual_camera(const emvs_camera& cam, const emvs_shape& shape, float out[4]);
bool host_pose_at(const emvs_stamped_pose* traj, size_t n, uint32_t sec, uint32_t nsec, emvs_pose* out);
void host_pose_compose(const emvs_pose& a, const emvs_pose& b, emvs_pose* out);
void host_pose_inverse(const emvs_pose& a, emvs_pose* out);
int host_rectify_lut(int model, const double K[9], const double* D, int n_d, const double R[9], const double P[12],
                     uint32_t W, uint32_t H, float* out);
// Where the packet stage reads an event's timestamp (the only event field the host touches): the 16-byte
// dvs_msgs::Event structs, or the int64 nanosecond array of an emvs_events_soa.
struct EventTimes {
  const emvs_event* aos = nullptr;
  const int64_t* t_ns = nullptr;
  void at(size_t i, uint32_t* sec, uint32_t* nsec) const
  {
    if (aos) {
      *sec = aos[i].sec;
      *nsec = aos[i].nsec;
    } else {
      const int64_t t = t_ns[i];
      *sec = (uint32_t)(t / 1000000000ll);
      *nsec = (uint32_t)(t % 1000000000ll);
    }
  }
};
inline EventTimes times_of(const emvs_event* ev) { EventTimes t; t.aos = ev; return t; }
// *truncated (may be NULL) is set when the loop stopped because `max_out` packets were written while more would fit.
size_t host_packetize(const EventTimes& ev, size_t n_ev, const emvs_stamped_pose* traj, size_t n_poses,
                      const emvs_pose& T_rv_w, const emvs_camera& cam, const float virt[4], float z0,
                      emvs_packet* out, size_t max_out, bool* truncated = nullptr);
size_t host_packetize_range(const EventTimes& ev, size_t n_ev, const emvs_stamped_pose* traj, size_t n_poses,
                            const emvs_pose& T_rv_w, const emvs_camera& cam, const float virt[4], float z0,
                            size_t* cur_inout, size_t event_limit, emvs_packet* out, size_t max_out,
                            bool* truncated = nullptr);

// nccl_dl.cpp — NCCL resolved at run time so the library loads (and every single-GPU entry
// point works) on machines without NCCL, and shares torch's NCCL when loaded from Python.
struct NcclId {  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES == 128), passed by value
  char b[128];
};
struct NcclApi {
  int (*GetUniqueId)(NcclId* id);
  int (*CommInitRank)(void** comm, int nranks, NcclId id, int rank);
  int (*CommDestroy)(void* comm);
  int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void* stream);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
  const char* (*GetErrorString)(int);
};
const NcclApi* nccl_api();  // nullptr when libnccl cannot be loaded (error set)

}  // namespace emvs

#endif
