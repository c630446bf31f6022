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
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace cup2d {

struct PullArgs {
  const double *peer[3][MAX_RANKS]; // base of the same buffer on every rank; [k]: candidate k when the buffer is chosen on the device
};

// `sel`: null, or the device-side index (KrylovState::opt) of the candidate buffer to exchange — the best Krylov iterate
// lives in one of three x buffers and which one is known only on the device when nothing synchronises with the host.
// src_mask / dst_mask: ranks this rank pulls from / ranks that pull from this rank.  Only those are signalled and waited
// for, so a halo refresh couples a rank to its (at most a handful of) neighbours, not to all eight GPUs.
__global__ void __launch_bounds__(256)
halo_pull_kernel(double *dst0, double *dst1, double *dst2, PullArgs pa, Comm comm, const int2 *__restrict__ src,
                 int nhalo, int nloc, int blk_doubles, const int *__restrict__ sel, unsigned src_mask, unsigned dst_mask,
                 unsigned int *counter, int done_barrier) {
  __shared__ bool s_last;
  unsigned long long *mine = comm.mb[comm.rank];
  const int tid = threadIdx.x;
  // the epoch of this refresh: every CTA reads the counter, the CTA that finishes last advances it
  const unsigned long long epoch = ld_relaxed_sys(mine + MB_HEPOCH) + 1;
  if (blockIdx.x == 0 && tid < comm.nranks && ((dst_mask >> tid) & 1u)) {
    __threadfence_system();
    st_release_sys(comm.mb[tid] + MB_READY + comm.rank, epoch);
  }
  if (tid < comm.nranks && ((src_mask >> tid) & 1u)) wait_flag(mine + MB_READY + tid, epoch, comm, CW_HALO_READY, tid);
  __syncthreads();
  const int which = sel ? *sel : 0;
  double *dst = which == 0 ? dst0 : (which == 1 ? dst1 : dst2);
  // one warp per halo block; 128-bit loads over NVLink
  const int warps = (gridDim.x * blockDim.x) >> 5, w = (blockIdx.x * blockDim.x + tid) >> 5, lane = tid & 31;
  const int n2 = blk_doubles >> 1; // double2 per block
  for (int k = w; k < nhalo; k += warps) {
    const int2 so = src[k];
    const double2 *from = reinterpret_cast<const double2 *>(pa.peer[which][so.x]) + (size_t)so.y * n2;
    double2 *to = reinterpret_cast<double2 *>(dst) + (size_t)(nloc + k) * n2;
    for (int i = lane; i < n2; i += 32) to[i] = ld_coherent2(from + i);
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // done_barrier = 0: the caller guarantees that a later all-reduce orders the owners' next write after these reads
  if (done_barrier && tid < comm.nranks) {
    if ((src_mask >> tid) & 1u) {
      __threadfence_system();
      st_release_sys(comm.mb[tid] + MB_DONE + comm.rank, epoch);
    }
    if ((dst_mask >> tid) & 1u) wait_flag(mine + MB_DONE + tid, epoch, comm, CW_HALO_DONE, tid);
  }
  __syncthreads();
  if (tid == 0) {
    *counter = 0;
    st_relaxed_sys(mine + MB_HEPOCH, epoch);
  }
}

static int halo_launch(cup2d_sim *s, double *const dst[3], const int peer_index[3], int ncand, const int *d_sel, int dim,
                       bool done_barrier) {
  if (s->nranks == 1) return CUP2D_OK;
  if (!s->peers_attached) {
    set_error("halo exchange before cup2d_peer_attach");
    return CUP2D_ESTATE;
  }
  PullArgs pa;
  memset(&pa, 0, sizeof pa);
  for (int k = 0; k < ncand; k++) {
    for (int r = 0; r < s->nranks; r++) pa.peer[k][r] = (const double *)s->peer_base[r][peer_index[k]];
    pa.peer[k][s->rank] = dst[k];
  }
  int grid = (int)((s->nhalo * 32 + 255) / 256);
  if (grid > s->num_sms) grid = s->num_sms;
  if (grid < 1) grid = 1;
  ProfScope prof(s, KC_HALO);
  halo_pull_kernel<<<grid, 256, 0, s->stream>>>(dst[0], ncand > 1 ? dst[1] : dst[0], ncand > 2 ? dst[2] : dst[0], pa, s->comm,
                                                reinterpret_cast<const int2 *>(s->d_halo_src), (int)s->nhalo, (int)s->nloc,
                                                64 * dim, d_sel, s->src_mask, s->dst_mask, s->d_counter, done_barrier ? 1 : 0);
  s->launches++;
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

int halo_exchange_ptr(cup2d_sim *s, double *base, int dim, int peer_index, bool done_barrier) {
  double *const dst[3] = {base, base, base};
  const int pi[3] = {peer_index, peer_index, peer_index};
  return halo_launch(s, dst, pi, 1, nullptr, dim, done_barrier);
}

// halo of the Krylov x buffer that holds the best iterate, chosen on the device by KrylovState::opt
int halo_exchange_xopt(cup2d_sim *s) {
  double *const dst[3] = {s->kx[0], s->kx[1], s->kx[2]};
  const int pi[3] = {CUP2D_NFIELDS + 1, CUP2D_NFIELDS + 2, CUP2D_NFIELDS + 3};
  return halo_launch(s, dst, pi, 3, &s->d_state->opt, 1, true);
}

// Error word of this rank's mailbox (first cross-GPU wait that was given up); CUP2D_ECOMM if set.  Called by the entry
// points that synchronise with the device anyway.
int comm_check(cup2d_sim *s) {
  if (s->nranks == 1 || !s->peers_attached) return CUP2D_OK;
  unsigned long long w = 0;
  CUP2D_CUDA(cudaMemcpy(&w, s->d_mailbox + MB_ERR, sizeof w, cudaMemcpyDeviceToHost));
  if (w == 0) return CUP2D_OK;
  static const char *what[] = {"?", "all-reduce contribution", "halo READY flag", "halo DONE flag", "pushed-halo flag"};
  const int wi = (int)((w >> 8) & 0xff);
  set_error("cross-GPU wait timed out on rank " + std::to_string(s->rank) + ": " + what[wi >= 1 && wi <= 4 ? wi : 0] +
            " of rank " + std::to_string((int)(w & 0xff)) + " never arrived (peer dead, or the ranks issued different call sequences)");
  return CUP2D_ECOMM;
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
  CUP2D_CUDA(cudaIpcGetMemHandle(&b.halo_gid, s->d_halo_gid));
  b.nloc = s->nloc;
  b.nhalo = s->nhalo;
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
  // Who pulls from whom.  src_mask: owners of this rank's halo slots.  The other direction — which of this rank's blocks
  // are halo slots of which peer, and where — comes out of the peers' own halo lists (read through the mapping): it gives
  // dst_mask for the pull kernel's flags and the push table of the Krylov kernels (poisson.cu: the producer of z writes
  // perimeter rows straight into the neighbours' halo slots).
  s->src_mask = s->dst_mask = 0;
  for (int64_t k = 0; k < s->nhalo; k++) s->src_mask |= 1u << s->halo_owner[k];
  std::vector<std::vector<int2>> per_block(s->nloc);
  int64_t nent = 0;
  for (int r = 0; r < s->nranks; r++) {
    if (r == s->rank || blobs[r].nhalo == 0) continue;
    void *p = nullptr;
    CUP2D_CUDA(cudaIpcOpenMemHandle(&p, blobs[r].halo_gid, cudaIpcMemLazyEnablePeerAccess));
    std::vector<int32_t> gid((size_t)blobs[r].nhalo);
    const cudaError_t e = cudaMemcpy(gid.data(), p, gid.size() * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaIpcCloseMemHandle(p);
    CUP2D_CUDA(e);
    for (int64_t k = 0; k < blobs[r].nhalo; k++) {
      const int64_t g = gid[k];
      if (g < s->gbegin || g >= s->gbegin + s->nloc) continue;
      per_block[g - s->gbegin].push_back(make_int2(r, (int)(blobs[r].nloc + k)));
      s->dst_mask |= 1u << r;
      nent++;
    }
  }
  std::vector<int> first((size_t)s->nloc, -1);
  std::vector<int2> ent;
  ent.reserve((size_t)nent + s->nloc / 8 + 1);
  for (int64_t b = 0; b < s->nloc; b++) {
    if (per_block[b].empty()) continue;
    first[b] = (int)ent.size();
    for (const int2 &e : per_block[b]) ent.push_back(e);
    ent.push_back(make_int2(-1, -1)); // terminator
  }
  if (ent.empty()) ent.push_back(make_int2(-1, -1));
  cudaFree(s->d_push_first);
  cudaFree(s->d_push_ent);
  s->d_push_first = nullptr;
  s->d_push_ent = nullptr;
  CUP2D_CUDA(cudaMalloc(&s->d_push_first, first.size() * sizeof(int) + 16));
  CUP2D_CUDA(cudaMemcpy(s->d_push_first, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUP2D_CUDA(cudaMalloc(&s->d_push_ent, ent.size() * sizeof(int2)));
  CUP2D_CUDA(cudaMemcpy(s->d_push_ent, ent.data(), ent.size() * sizeof(int2), cudaMemcpyHostToDevice));
  s->n_push = nent;
  if (const char *t = getenv("CUP2D_COMM_TIMEOUT_MS")) s->comm.timeout_ns = (unsigned long long)(atof(t) * 1e6);
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
