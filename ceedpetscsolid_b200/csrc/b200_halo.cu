// Halo exchange over NVLink peer memory for the /gpu/b200 harness (sm_100a).
//
// DMLocalToGlobal(ADD_VALUES) + DMGlobalToLocal(INSERT_VALUES) of the reference (/root/reference/src/matops.c:33,57)
// as ONE symmetric sum-and-share between the ranks of a node, without a communication library on the data path:
//
//   k_halo_push         every shared dof's partial sum is STORED straight into the receive window of each
//                       neighbour that holds it (peer pointers obtained once through CUDA IPC), system fence;
//   k_halo_signal       one release-store per neighbour: "generation g of my data is in your window";
//   k_halo_wait_unpack  spins (bounded) until every neighbour's flag shows generation g, then adds the window
//                       into the L-vector.
//
// Windows are double-buffered by generation parity: a neighbour can be at most one exchange ahead (it cannot
// finish exchange g+1 without my push g+1, which is stream-ordered after my unpack g), so parity g+1 is free
// when it writes.  A spin that exceeds its cycle budget sets an error word and lets the kernel finish: a lost
// peer makes the run fail loudly instead of hanging the GPU.
#include <string.h>

#include "b200_common.cuh"

namespace b200 {

struct HaloSegs {
  int n;
  int start[B200_HALO_MAX_NEIGHBOURS + 1];          // segment s covers packed positions [start[s], start[s+1])
  double *remote[B200_HALO_MAX_NEIGHBOURS];         // where segment s goes in neighbour s's window (this parity)
  long long *remote_flag[B200_HALO_MAX_NEIGHBOURS]; // my slot in neighbour s's flag array
};

__global__ void k_halo_push(const __grid_constant__ HaloSegs sg, const int *__restrict__ idx,
                            const double *__restrict__ y) {
  const size_t total = (size_t)sg.start[sg.n];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int s = 0;
    while (s + 1 < sg.n && i >= (size_t)sg.start[s + 1]) s++;
    sg.remote[s][i - sg.start[s]] = y[idx[i]];
  }
  __threadfence_system();
}

__global__ void k_halo_signal(const __grid_constant__ HaloSegs sg, long long gen) {
  const int s = threadIdx.x;
  if (s < sg.n) {
    __threadfence_system();
    *reinterpret_cast<volatile long long *>(sg.remote_flag[s]) = gen;
    __threadfence_system();
  }
}

__global__ void k_halo_wait_unpack(const long long *flags, int nnbr, long long gen, const int *__restrict__ idx,
                                   const double *window, double *__restrict__ y, size_t total, int *err,
                                   long long budget_cycles) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if (threadIdx.x < nnbr) {
    const volatile long long *f = flags + threadIdx.x;
    const long long t0 = clock64();
    while (*f < gen) {
      if (clock64() - t0 > budget_cycles) {
        bad = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (bad) {
    if (threadIdx.x == 0) atomicExch(err, 1);
    return;
  }
  __threadfence_system();
  // the window was written by the peers through NVLink into this device's memory: read it at L2 (no stale L1 lines)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(y + idx[i], __ldcg(window + i));
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_ipc_get_handle(const void *dptr, unsigned char *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  B200_CHECK(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
  memcpy(handle64, &h, 64);
  return 0;
}

int b200_ipc_open(const unsigned char *handle64, void **dptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B200_CHECK(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int b200_ipc_close(void *dptr) {
  if (dptr) B200_CHECK(cudaIpcCloseMemHandle(dptr));
  return 0;
}

static int fill_segs(HaloSegs &sg, int nnbr, const int *seg_start, double *const *remote, long long *const *remote_flag) {
  if (nnbr < 0 || nnbr > B200_HALO_MAX_NEIGHBOURS) return set_error_msg("halo: too many neighbours");
  sg.n = nnbr;
  for (int s = 0; s <= nnbr; s++) sg.start[s] = seg_start[s];
  for (int s = 0; s < nnbr; s++) {
    sg.remote[s] = remote ? remote[s] : nullptr;
    sg.remote_flag[s] = remote_flag ? remote_flag[s] : nullptr;
  }
  return 0;
}

int b200_halo_push_signal(int nnbr, const int *seg_start, double *const *remote, long long *const *remote_flag,
                          const int *d_idx, const double *d_y, long long gen) {
  if (nnbr == 0) return 0;
  HaloSegs sg;
  if (int rc = fill_segs(sg, nnbr, seg_start, remote, remote_flag)) return rc;
  const size_t total = (size_t)seg_start[nnbr];
  if (total) {
    size_t nb = (total + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    k_halo_push<<<(unsigned)nb, 256, 0, g_stream>>>(sg, d_idx, d_y);
    B200_LAUNCH_CHECK("k_halo_push");
  }
  k_halo_signal<<<1, 32, 0, g_stream>>>(sg, gen);
  B200_LAUNCH_CHECK("k_halo_signal");
  return 0;
}

int b200_halo_wait_unpack(int nnbr, const long long *d_flags, long long gen, const int *d_idx, const double *d_window,
                          double *d_y, size_t total, int *d_err, double timeout_s) {
  if (nnbr == 0) return 0;
  if (nnbr > B200_HALO_MAX_NEIGHBOURS) return set_error_msg("halo: too many neighbours");
  size_t nb = (total + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb == 0) nb = 1;
  const long long budget = (long long)(timeout_s * 1.9e9);
  k_halo_wait_unpack<<<(unsigned)nb, 256, 0, g_stream>>>(d_flags, nnbr, gen, d_idx, d_window, d_y, total, d_err, budget);
  B200_LAUNCH_CHECK("k_halo_wait_unpack");
  return 0;
}

}  // extern "C"
