// Inter-GPU halo exchange over NVLink peer memory (replaces sync1/Setup/pack/unpack_subregion +
// MPI_Isend/Irecv on host buffers, main.cpp:1971-2142, 909-1380, 58-110, and the host-staged halo of
// the Krylov operand, cuda.cu:344-380).
//
// One process per GPU.  Each rank cudaMalloc's its fields, exports CUDA IPC handles
// (cup2d_peer_export); the launcher all-gathers the blobs (torch.distributed is only plumbing) and
// every rank maps its peers' buffers (cup2d_peer_attach).  A halo refresh is ONE kernel that
//   1. publishes "my data of this epoch is ready" to every peer's mailbox (st.release.sys),
//   2. waits until every peer is ready (ld.acquire.sys on the local mailbox),
//   3. PULLS the face-neighbour blocks straight out of the owners' field arrays with 128-bit peer
//      loads into the local halo slots (whole 8x8 blocks: 512 B / 1 KB, so the stencil kernels see
//      remote neighbours exactly like local ones),
//   4. publishes "done reading" and waits for every peer's "done", so that when the kernel retires
//      the owners may overwrite the buffer (WAR safety without any host involvement).  Inside the
//      Krylov loop step 4 is skipped: the all-reduce fused into the SpMV kernel that consumes the halo
//      completes only after every rank finished its pull, and the owner's next write comes after it.
// No pack/unpack buffers, no host staging, no NCCL call on the critical path.  Payload at 8192^2 on
// 8 GPUs: ~1500 perimeter blocks -> 1.5 MB (velocity) / 0.75 MB (scalar) per refresh: latency-bound.
#include "sim.h"

namespace cup2d {

struct PullArgs {
  const double *peer[MAX_RANKS]; // base of the same buffer on every rank (peer[rank] unused)
};

__global__ void __launch_bounds__(256)
halo_pull_kernel(double *__restrict__ dst, PullArgs pa, Comm comm, const int2 *__restrict__ src,
                 int nhalo, int nloc, int blk_doubles, unsigned long long epoch,
                 unsigned int *counter, int done_barrier) {
  __shared__ bool s_last;
  unsigned long long *mine = comm.mb[comm.rank];
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid < comm.nranks && tid != comm.rank) {
    __threadfence_system();
    st_release_sys(comm.mb[tid] + MB_READY + comm.rank, epoch);
  }
  if (tid < comm.nranks && tid != comm.rank) {
    while (ld_acquire_sys(mine + MB_READY + tid) < epoch) { }
  }
  __syncthreads();
  // one warp per halo block; 128-bit loads over NVLink
  const int warps = (gridDim.x * blockDim.x) >> 5, w = (blockIdx.x * blockDim.x + tid) >> 5, lane = tid & 31;
  const int n2 = blk_doubles >> 1; // double2 per block
  for (int k = w; k < nhalo; k += warps) {
    const int2 so = src[k];
    const double2 *from = reinterpret_cast<const double2 *>(pa.peer[so.x]) + (size_t)so.y * n2;
    double2 *to = reinterpret_cast<double2 *>(dst) + (size_t)(nloc + k) * n2;
    for (int i = lane; i < n2; i += 32) {
      double2 v;
      asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(from + i) : "memory");
      to[i] = v;
    }
  }
  if (!done_barrier) return; // the caller guarantees a later all-reduce orders the owners' next write
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (tid < comm.nranks && tid != comm.rank) {
    __threadfence_system();
    st_release_sys(comm.mb[tid] + MB_DONE + comm.rank, epoch);
    while (ld_acquire_sys(mine + MB_DONE + tid) < epoch) { }
  }
  if (tid == 0) *counter = 0;
}

int halo_exchange_ptr(cup2d_sim *s, double *base, int dim, int peer_index, bool done_barrier) {
  if (s->nranks == 1) return CUP2D_OK;
  if (!s->peers_attached) {
    set_error("halo exchange before cup2d_peer_attach");
    return CUP2D_ESTATE;
  }
  PullArgs pa;
  for (int r = 0; r < MAX_RANKS; r++) pa.peer[r] = r < s->nranks ? (const double *)s->peer_base[r][peer_index] : nullptr;
  pa.peer[s->rank] = base;
  s->epoch++;
  int grid = (int)((s->nhalo * 32 + 255) / 256);
  if (grid > s->num_sms) grid = s->num_sms;
  if (grid < 1) grid = 1;
  ProfScope prof(s, KC_HALO);
  halo_pull_kernel<<<grid, 256, 0, s->stream>>>(base, pa, s->comm, reinterpret_cast<const int2 *>(s->d_halo_src),
                                                (int)s->nhalo, (int)s->nloc, 64 * dim, s->epoch, s->d_counter, done_barrier ? 1 : 0);
  s->launches++;
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

void swap_fields(cup2d_sim *s, int a, int b) {
  std::swap(s->f[a], s->f[b]);
  for (int r = 0; r < s->nranks; r++) std::swap(s->peer_base[r][a], s->peer_base[r][b]);
}

} // namespace cup2d

using namespace cup2d;

extern "C" {

int cup2d_peer_blob_size(void) { return (int)sizeof(PeerBlob); }

int cup2d_peer_export(cup2d_sim *s, void *blob) {
  if (!s || !blob) {
    set_error("cup2d_peer_export: null argument");
    return CUP2D_EINVAL;
  }
  CUP2D_CUDA(cudaSetDevice(s->device));
  PeerBlob b;
  memset(&b, 0, sizeof b);
  for (int f = 0; f < CUP2D_NFIELDS; f++) CUP2D_CUDA(cudaIpcGetMemHandle(&b.field[f], s->f[f]));
  CUP2D_CUDA(cudaIpcGetMemHandle(&b.kz, s->kz));
  for (int k = 0; k < 3; k++) CUP2D_CUDA(cudaIpcGetMemHandle(&b.kx[k], s->kx[k]));
  CUP2D_CUDA(cudaIpcGetMemHandle(&b.kzr, s->kzr));
  CUP2D_CUDA(cudaIpcGetMemHandle(&b.mailbox, s->d_mailbox));
  b.nloc = s->nloc;
  b.rank = s->rank;
  b.device = s->device;
  memcpy(blob, &b, sizeof b);
  return CUP2D_OK;
}

int cup2d_peer_attach(cup2d_sim *s, const void *all_blobs) {
  if (!s || !all_blobs) {
    set_error("cup2d_peer_attach: null argument");
    return CUP2D_EINVAL;
  }
  if (s->nranks == 1) {
    s->peers_attached = true;
    return CUP2D_OK;
  }
  CUP2D_CUDA(cudaSetDevice(s->device));
  const PeerBlob *blobs = static_cast<const PeerBlob *>(all_blobs);
  for (int r = 0; r < s->nranks; r++) {
    if (blobs[r].rank != r) {
      set_error("cup2d_peer_attach: blobs are not in rank order");
      return CUP2D_EINVAL;
    }
    if (r == s->rank) {
      for (int f = 0; f < CUP2D_NFIELDS; f++) s->peer_base[r][f] = s->f[f];
      s->peer_base[r][CUP2D_NFIELDS] = s->kz;
      for (int k = 0; k < 3; k++) s->peer_base[r][CUP2D_NFIELDS + 1 + k] = s->kx[k];
      s->peer_base[r][CUP2D_NFIELDS + 4] = s->kzr;
      s->peer_mailbox[r] = s->d_mailbox;
      continue;
    }
    for (int f = 0; f < CUP2D_NFIELDS; f++)
      CUP2D_CUDA(cudaIpcOpenMemHandle(&s->peer_base[r][f], blobs[r].field[f], cudaIpcMemLazyEnablePeerAccess));
    CUP2D_CUDA(cudaIpcOpenMemHandle(&s->peer_base[r][CUP2D_NFIELDS], blobs[r].kz, cudaIpcMemLazyEnablePeerAccess));
    for (int k = 0; k < 3; k++)
      CUP2D_CUDA(cudaIpcOpenMemHandle(&s->peer_base[r][CUP2D_NFIELDS + 1 + k], blobs[r].kx[k], cudaIpcMemLazyEnablePeerAccess));
    CUP2D_CUDA(cudaIpcOpenMemHandle(&s->peer_base[r][CUP2D_NFIELDS + 4], blobs[r].kzr, cudaIpcMemLazyEnablePeerAccess));
    void *mb = nullptr;
    CUP2D_CUDA(cudaIpcOpenMemHandle(&mb, blobs[r].mailbox, cudaIpcMemLazyEnablePeerAccess));
    s->peer_mailbox[r] = static_cast<unsigned long long *>(mb);
  }
  s->comm.rank = s->rank;
  s->comm.nranks = s->nranks;
  for (int r = 0; r < s->nranks; r++) s->comm.mb[r] = s->peer_mailbox[r];
  s->peers_attached = true;
  return CUP2D_OK;
}

/* device address of `field` on rank `rank` as mapped into this process (this rank: its own array); null before attach */
void *cup2d_peer_field_ptr(cup2d_sim *s, int rank, int field) {
  if (!s || rank < 0 || rank >= s->nranks || field < 0 || field >= CUP2D_NFIELDS || !s->peers_attached) return nullptr;
  return s->nranks == 1 ? (void *)s->f[field] : s->peer_base[rank][field];
}

int cup2d_halo_exchange(cup2d_sim *s, int field) {
  if (!s || field < 0 || field >= CUP2D_NFIELDS) {
    set_error("cup2d_halo_exchange: bad argument");
    return CUP2D_EINVAL;
  }
  CUP2D_CUDA(cudaSetDevice(s->device));
  return halo_exchange_ptr(s, s->f[field], dim_of(field), field);
}

} // extern "C"
