// halo_p2p.cu -- ghost-cell exchange between ranks of one NVLink / NVSwitch box by peer-mapped stores.
//
// Replaces, for ranks that share a box, the reference's exchange path: Load_Hydro_DeviceBuffer_* -> PackBuffers3D ->
// MPI_Isend / MPI_Irecv of the packed slab -> MPI_Waitany -> UnpackBuffers3D (src/grid/mpi_boundaries.cpp:113-298,
// 300-830, src/grid/cuda_boundaries.cu:11-69), which stages every slab through a send buffer, the MPI library and a
// receive buffer with a host wait per message.  Here the pack kernel of the SENDER stores the slab straight into the
// RECEIVER's buffer over NVLink (the receiver's block is mapped into the sender's address space with CUDA IPC), and
// raises an arrival counter in the receiver's memory once every block's stores are fenced; the receiver's unpack
// kernel waits on that counter on the device.  No host wait, no library call and no intermediate copy on the data
// path: per exchanged face one send launch and one receive launch, all stream-ordered.
//
// Ordering rules that make this safe without any extra handshake:
//   * buffers are double-buffered by the parity of the exchange number: a sender can only be two exchanges ahead of a
//     receiver's unpack if it has already consumed that receiver's next message, which the receiver sends after the
//     unpack in question (stream order) -- so writing exchange s+2 into buffer (s mod 2) never overtakes its reader;
//   * a rank's sends of an axis precede its receives of that axis in its stream, so two ranks waiting on each other
//     cannot deadlock; a receive that waits longer than kTimeoutNs gives up and raises the error word instead of
//     hanging the GPU (chl_halo_error).
// The x -> y -> z phase order, with the y slabs carrying the x ghosts and the z slabs the x and y ghosts, is the
// reference's (mpi_boundaries.cpp:29-111): edges and corners need no diagonal messages.
#include <stdint.h>

#include "chl_aux.h"

namespace chl
{
namespace
{
constexpr unsigned long long kTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// slab of the local array -> the neighbour's receive buffer [field][k][j][i] (the layout of PackBuffers3D); the last
// block to finish publishes the exchange number in the neighbour's arrival word
template <typename T>
__global__ void __launch_bounds__(256) halo_send_kernel(T *__restrict__ remote, const T *__restrict__ c, int nx, int ny,
                                                        int n_fields, long n_cells, long idxoffset, int isize, int jsize,
                                                        int ksize, unsigned *counter, unsigned long long *remote_flag,
                                                        unsigned long long seq)
{
  long const nb = (long)isize * jsize * ksize;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < nb; t += (long)gridDim.x * blockDim.x) {
    int const i  = (int)(t % isize);
    long const r = t / isize;
    int const j = (int)(r % jsize), k = (int)(r / jsize);
    long const idx = i + ((long)j + (long)k * ny) * nx + idxoffset;
    for (int f = 0; f < n_fields; f++) remote[t + f * nb] = c[idx + f * n_cells];
  }
  __threadfence_system();  // this thread's stores are visible to the peer before the block reports in
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned const done = atomicAdd(counter, 1u) + 1u;
    if (done == gridDim.x) {
      *counter = 0u;  // ready for the next launch on this face (stream-ordered)
      __threadfence_system();
      st_release_sys(remote_flag, seq);
    }
  }
}

// wait for exchange `seq` to have arrived in this rank's buffer, then buffer -> ghost slab (UnpackBuffers3D)
template <typename T>
__global__ void __launch_bounds__(256) halo_recv_kernel(const T *__restrict__ buf, T *__restrict__ c, int nx, int ny,
                                                        int n_fields, long n_cells, long idxoffset, int isize, int jsize,
                                                        int ksize, const unsigned long long *flag, unsigned long long seq,
                                                        unsigned *err)
{
  __shared__ int ok;
  if (threadIdx.x == 0) {
    unsigned long long const t0 = global_ns();
    int good = 1;
    while (ld_acquire_sys(flag) < seq) {
      if (global_ns() - t0 > kTimeoutNs) {
        good = 0;
        atomicExch(err, 1u);
        break;
      }
      __nanosleep(200);
    }
    ok = good;
  }
  __syncthreads();
  if (!ok) return;
  long const nb = (long)isize * jsize * ksize;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < nb; t += (long)gridDim.x * blockDim.x) {
    int const i  = (int)(t % isize);
    long const r = t / isize;
    int const j = (int)(r % jsize), k = (int)(r / jsize);
    long const idx = i + ((long)j + (long)k * ny) * nx + idxoffset;
    for (int f = 0; f < n_fields; f++) c[idx + f * n_cells] = __ldcg(buf + t + f * nb);  // L2: written by the peer
  }
}

inline int grid_for(long n)
{
  long b = (n + 255) / 256;
  if (b > 148 * 4) b = 148 * 4;  // a copy at NVLink speed needs few CTAs; the interior update keeps the rest of the SMs
  if (b < 1) b = 1;
  return (int)b;
}
}  // namespace

cudaError_t launch_halo_send(void *remote, const void *c, int nx, int ny, int n_fields, long n_cells, long idxoffset,
                             int isize, int jsize, int ksize, unsigned *counter, unsigned long long *remote_flag,
                             unsigned long long seq, size_t real_bytes, cudaStream_t s)
{
  long const nb = (long)isize * jsize * ksize;
  if (nb <= 0) return cudaErrorInvalidValue;
  int const g = grid_for(nb);
  if (real_bytes == 8)
    halo_send_kernel<double><<<g, 256, 0, s>>>((double *)remote, (const double *)c, nx, ny, n_fields, n_cells, idxoffset,
                                                isize, jsize, ksize, counter, remote_flag, seq);
  else
    halo_send_kernel<float><<<g, 256, 0, s>>>((float *)remote, (const float *)c, nx, ny, n_fields, n_cells, idxoffset, isize,
                                               jsize, ksize, counter, remote_flag, seq);
  return cudaGetLastError();
}

cudaError_t launch_halo_recv(const void *buf, void *c, int nx, int ny, int n_fields, long n_cells, long idxoffset, int isize,
                             int jsize, int ksize, const unsigned long long *flag, unsigned long long seq, unsigned *err,
                             size_t real_bytes, cudaStream_t s)
{
  long const nb = (long)isize * jsize * ksize;
  if (nb <= 0) return cudaErrorInvalidValue;
  int const g = grid_for(nb);
  if (real_bytes == 8)
    halo_recv_kernel<double><<<g, 256, 0, s>>>((const double *)buf, (double *)c, nx, ny, n_fields, n_cells, idxoffset, isize,
                                                jsize, ksize, flag, seq, err);
  else
    halo_recv_kernel<float><<<g, 256, 0, s>>>((const float *)buf, (float *)c, nx, ny, n_fields, n_cells, idxoffset, isize,
                                               jsize, ksize, flag, seq, err);
  return cudaGetLastError();
}
}  // namespace chl
