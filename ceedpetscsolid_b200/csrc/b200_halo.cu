// Halo exchange over NVLink peer memory for the /gpu/b200 harness (sm_100a).
//
// DMLocalToGlobal(ADD_VALUES) + DMGlobalToLocal(INSERT_VALUES) of the reference (/root/reference/src/matops.c:33,57)
// as ONE symmetric sum-and-share between the ranks of a node, without a communication library on the data path:
//
//   k_halo_push         every shared dof's partial sum is STORED straight into the receive window of each
//                       neighbour that holds it (peer pointers obtained once through CUDA IPC), system fence;
//   k_halo_signal       one release-store per neighbour: "generation g of my data is in your window";
//   k_halo_wait_unpack  spins (bounded) until every neighbour's flag shows generation g, then, for every shared
//                       dof, sums the partial sums of ALL its holders in ascending rank order (its own included):
//                       every holder ends up with the bit-identical value, no atomics.
//
// b200_halo_begin runs push + signal on a high-priority side stream behind an event of the compute stream, so a
// caller that has finished the elements touching the partition interface can keep the compute stream busy with the
// interior elements while the data crosses NVLink; b200_halo_end queues the wait + unpack on the compute stream.
// The generation counter lives in device memory and is advanced by the signal kernel: no per-call host state, the
// whole sequence can be captured in a CUDA graph.
//
// Windows are double-buffered by generation parity: a neighbour can be at most one exchange ahead (it cannot
// finish exchange g+1 without my push g+1, which is ordered after my unpack g), so parity g+1 is free when it
// writes.  A spin that exceeds its cycle budget sets an error word and lets the kernel finish: a lost peer makes
// the run fail loudly (b200_halo_error) instead of hanging the GPU.
#include <stdlib.h>
#include <string.h>

#include "b200_common.cuh"

namespace b200 {

struct HaloSegs {
  int n;
  int start[B200_HALO_MAX_NEIGHBOURS + 1];            // segment s covers packed positions [start[s], start[s+1])
  double *remote[2][B200_HALO_MAX_NEIGHBOURS];        // where segment s goes in neighbour s's window, per parity
  long long *remote_flag[B200_HALO_MAX_NEIGHBOURS];   // my slot in neighbour s's flag array
};

__global__ void k_halo_push(const __grid_constant__ HaloSegs sg, const int *__restrict__ idx,
                            const double *__restrict__ y, const long long *__restrict__ gen) {
  const int par = (int)((*gen + 1) & 1);
  const size_t total = (size_t)sg.start[sg.n];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int s = 0;
    while (s + 1 < sg.n && i >= (size_t)sg.start[s + 1]) s++;
    sg.remote[par][s][i - sg.start[s]] = y[idx[i]];
  }
  __threadfence_system();
}

__global__ void k_halo_signal(const __grid_constant__ HaloSegs sg, long long *gen) {
  const long long g = *gen + 1;
  const int s = threadIdx.x;
  if (s < sg.n) {
    __threadfence_system();
    *reinterpret_cast<volatile long long *>(sg.remote_flag[s]) = g;
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *gen = g;
}

// sum over the holders of unique shared dof u, ascending rank: entry < 0 is this rank's own partial sum (in y),
// entry >= 0 a position of the receive window
__device__ __forceinline__ void ordered_sum(int u, const int *__restrict__ udof, const int *__restrict__ uptr,
                                            const int *__restrict__ uent, const double *window, double *__restrict__ y) {
  const int d = udof[u];
  double s = 0.0;
  for (int j = uptr[u]; j < uptr[u + 1]; j++) {
    const int en = uent[j];
    s += en < 0 ? y[d] : __ldcg(window + en);   // the window was written by peers: read it at L2
  }
  y[d] = s;
}

__global__ void k_halo_wait_unpack(const long long *flags, int nnbr, const long long *gen, const double *window_base,
                                   size_t total, int nuniq, const int *__restrict__ udof, const int *__restrict__ uptr,
                                   const int *__restrict__ uent, double *__restrict__ y, int *err, long long budget_cycles) {
  __shared__ int bad;
  const long long g = *gen;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if (threadIdx.x < nnbr) {
    const volatile long long *f = flags + threadIdx.x;
    const long long t0 = clock64();
    while (*f < g) {
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
  const double *window = window_base + (size_t)(g & 1) * total;
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < (size_t)nuniq; u += (size_t)gridDim.x * blockDim.x)
    ordered_sum((int)u, udof, uptr, uent, window, y);
}

__global__ void k_halo_unpack_ordered(int nuniq, const int *__restrict__ udof, const int *__restrict__ uptr,
                                      const int *__restrict__ uent, const double *window, double *__restrict__ y) {
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < (size_t)nuniq; u += (size_t)gridDim.x * blockDim.x)
    ordered_sum((int)u, udof, uptr, uent, window, y);
}

}  // namespace b200

using namespace b200;

struct b200_halo {
  HaloSegs sg;
  const int *d_idx;
  size_t total;
  double *d_window;      // [parity 0 | parity 1]
  long long *d_flags;    // one per neighbour (B200_HALO_MAX_NEIGHBOURS slots)
  long long *d_gen;
  int *d_err;
  int nuniq;
  const int *d_udof, *d_uptr, *d_uent;
  long long budget_cycles;
  cudaStream_t side;
  cudaEvent_t ev_ready, ev_pushed;
  int device;
};

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

size_t b200_halo_window_bytes(size_t total) {
  return 2 * total * sizeof(double) + (B200_HALO_MAX_NEIGHBOURS + 1) * sizeof(long long) + 16;
}

int b200_halo_create(int nnbr, const int *seg_start, double *const *remote_p0, double *const *remote_p1,
                     long long *const *remote_flag, const int *d_idx, size_t total, void *d_window, int nuniq,
                     const int *d_udof, const int *d_uptr, const int *d_uent, double timeout_s, b200_halo **out) {
  if (nnbr < 0 || nnbr > B200_HALO_MAX_NEIGHBOURS) return set_error_msg("halo: too many neighbours");
  b200_halo *h = (b200_halo *)calloc(1, sizeof *h);
  if (!h) return set_error_msg("halo: out of memory");
  h->sg.n = nnbr;
  for (int s = 0; s <= nnbr; s++) h->sg.start[s] = seg_start[s];
  for (int s = 0; s < nnbr; s++) {
    h->sg.remote[0][s] = remote_p0[s];
    h->sg.remote[1][s] = remote_p1[s];
    h->sg.remote_flag[s] = remote_flag[s];
  }
  h->d_idx = d_idx;
  h->total = total;
  h->d_window = (double *)d_window;
  h->d_flags = (long long *)((char *)d_window + 2 * total * sizeof(double));
  h->d_gen = h->d_flags + B200_HALO_MAX_NEIGHBOURS;
  h->d_err = (int *)(h->d_gen + 1);
  h->nuniq = nuniq;
  h->d_udof = d_udof; h->d_uptr = d_uptr; h->d_uent = d_uent;
  B200_CHECK(cudaGetDevice(&h->device));
  int khz = 0;
  B200_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device));   // clock64 ticks at the SM clock
  h->budget_cycles = (long long)(timeout_s * 1e3 * (double)(khz > 0 ? khz : 1900000));
  int lo = 0, hi = 0;
  B200_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  B200_CHECK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));   // hi = numerically lowest = highest priority
  B200_CHECK(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  B200_CHECK(cudaEventCreateWithFlags(&h->ev_pushed, cudaEventDisableTiming));
  *out = h;
  return 0;
}

int b200_halo_destroy(b200_halo *h) {
  if (!h) return 0;
  cudaStreamSynchronize(h->side);
  cudaEventDestroy(h->ev_ready);
  cudaEventDestroy(h->ev_pushed);
  cudaStreamDestroy(h->side);
  free(h);
  return 0;
}

// Fork: the side stream waits for everything queued on the compute stream so far and is returned, so that the caller
// can queue the work that PRODUCES the interface partial sums on it (b200_set_stream(side) ... b200_set_stream(main))
// while the compute stream carries on with work that does not touch shared dofs.  b200_halo_begin_forked then pushes
// behind that work without another event.
int b200_halo_fork(b200_halo *h, void **side_stream) {
  B200_CHECK(cudaEventRecord(h->ev_ready, g_stream));
  B200_CHECK(cudaStreamWaitEvent(h->side, h->ev_ready, 0));
  *side_stream = (void *)h->side;
  return 0;
}

static int halo_push_on_side(b200_halo *h, const double *d_y) {
  if (h->total) {
    size_t nb = (h->total + 255) / 256;
    if (nb > 148 * 4) nb = 148 * 4;
    k_halo_push<<<(unsigned)nb, 256, 0, h->side>>>(h->sg, h->d_idx, d_y, h->d_gen);
    B200_LAUNCH_CHECK("k_halo_push");
  }
  k_halo_signal<<<1, 32, 0, h->side>>>(h->sg, h->d_gen);
  B200_LAUNCH_CHECK("k_halo_signal");
  B200_CHECK(cudaEventRecord(h->ev_pushed, h->side));
  return 0;
}

// after b200_halo_fork: the interface partial sums are produced by work already queued on the side stream
int b200_halo_begin_forked(b200_halo *h, const double *d_y) {
  if (h->sg.n == 0) {   // nothing to exchange, but the compute stream must still join the side stream's work
    B200_CHECK(cudaEventRecord(h->ev_pushed, h->side));
    return 0;
  }
  return halo_push_on_side(h, d_y);
}

// every partial sum on a shared dof must be complete on the compute stream when this is called
int b200_halo_begin(b200_halo *h, const double *d_y) {
  if (h->sg.n == 0) return 0;
  B200_CHECK(cudaEventRecord(h->ev_ready, g_stream));
  B200_CHECK(cudaStreamWaitEvent(h->side, h->ev_ready, 0));
  return halo_push_on_side(h, d_y);
}

int b200_halo_end(b200_halo *h, double *d_y) {
  if (h->sg.n == 0) {   // joins a forked side stream (no-op otherwise: the event is then already complete)
    B200_CHECK(cudaStreamWaitEvent(g_stream, h->ev_pushed, 0));
    return 0;
  }
  // my own partial sums must have left (push reads y) before the unpack overwrites them with the totals
  B200_CHECK(cudaStreamWaitEvent(g_stream, h->ev_pushed, 0));
  size_t nb = ((size_t)h->nuniq + 255) / 256;
  if (nb > 148 * 4) nb = 148 * 4;
  if (nb == 0) nb = 1;
  k_halo_wait_unpack<<<(unsigned)nb, 256, 0, g_stream>>>(h->d_flags, h->sg.n, h->d_gen, h->d_window, h->total, h->nuniq,
                                                        h->d_udof, h->d_uptr, h->d_uent, d_y, h->d_err, h->budget_cycles);
  B200_LAUNCH_CHECK("k_halo_wait_unpack");
  return 0;
}

// synchronises; *err != 0: an exchange timed out waiting for a neighbour (the vector it produced is incomplete)
int b200_halo_error(b200_halo *h, int *err) {
  B200_CHECK(cudaStreamSynchronize(h->side));
  B200_CHECK(cudaStreamSynchronize(g_stream));
  B200_CHECK(cudaMemcpy(err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

// ordered unpack for exchanges carried by a communication library (receive buffer laid out like the window)
int b200_halo_unpack_ordered(int nuniq, const int *d_udof, const int *d_uptr, const int *d_uent, const double *d_recv,
                             double *d_y) {
  if (nuniq <= 0) return 0;
  size_t nb = ((size_t)nuniq + 255) / 256;
  if (nb > 148 * 4) nb = 148 * 4;
  k_halo_unpack_ordered<<<(unsigned)nb, 256, 0, g_stream>>>(nuniq, d_udof, d_uptr, d_uent, d_recv, d_y);
  B200_LAUNCH_CHECK("k_halo_unpack_ordered");
  return 0;
}

}  // extern "C"
