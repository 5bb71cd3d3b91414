// dist.cu -- NCCL binding (run-time) and the collectives the sharded build uses.
#include "dist.h"

#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "ctx.h"

namespace bgx {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // by SONAME: if the process already holds an NCCL (e.g. the one torch bundles) that one is reused
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.handle = h;
#define BGX_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
    BGX_SYM(GetUniqueId, "ncclGetUniqueId");
    BGX_SYM(CommInitRank, "ncclCommInitRank");
    BGX_SYM(CommDestroy, "ncclCommDestroy");
    BGX_SYM(AllReduce, "ncclAllReduce");
    BGX_SYM(AllGather, "ncclAllGather");
    BGX_SYM(Send, "ncclSend");
    BGX_SYM(Recv, "ncclRecv");
    BGX_SYM(GroupStart, "ncclGroupStart");
    BGX_SYM(GroupEnd, "ncclGroupEnd");
    BGX_SYM(GetErrorString, "ncclGetErrorString");
#undef BGX_SYM
  });
  BGX_CHECK(api.handle && api.GetUniqueId && api.CommInitRank && api.Send && api.Recv && api.GroupStart && api.GroupEnd &&
                api.AllReduce && api.AllGather,
            "NCCL (libnccl.so.2) is not available: multi-GPU builds need it");
  return api;
}

#define BGX_NCCL(expr)                                                                                 \
  do {                                                                                                 \
    ncclResult_t r__ = (expr);                                                                         \
    if (r__ != ncclSuccess)                                                                            \
      throw ::bgx::Error(std::string(#expr) + " failed: " + (nccl().GetErrorString ? nccl().GetErrorString(r__) : "?")); \
  } while (0)

inline ncclComm_t comm_of(Context* c) { return reinterpret_cast<ncclComm_t>(c->dist.comm); }

}  // namespace

void dist_get_unique_id(uint8_t id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  BGX_NCCL(nccl().GetUniqueId(&u));
  memcpy(id, &u, 128);
}

void dist_init(Context* c, int nranks, int rank, const uint8_t id[128]) {
  BGX_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bgx_dist_init: bad rank / world size");
  BGX_CHECK((nranks & (nranks - 1)) == 0 && nranks <= 64, "bgx_dist_init: world size must be a power of two <= 64");
  BGX_CHECK(c->dist.comm == nullptr, "bgx_dist_init: already initialised");
  BGX_CHECK(c->n_reads == 0, "bgx_dist_init: call before adding reads");
  c->dist.nranks = nranks;
  c->dist.rank = rank;
  if (nranks == 1) return;
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t comm = nullptr;
  BGX_NCCL(nccl().CommInitRank(&comm, nranks, u, rank));
  c->dist.comm = comm;
}

void dist_map_peers_n(Context* c, void* const* local, int k, void** peer) {
  const int N = c->dist.nranks, R = c->dist.rank;
  for (int j = 0; j < k; ++j) peer[(size_t)j * N + R] = local[j];
  if (N == 1) return;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  // per block: IPC handle (8 words), exporting process, raw pointer, device ordinal.  Ranks that live
  // in THIS process (bgx_bs::multi_session: one host thread per GPU) cannot open their own handles;
  // their memory is addressed directly once peer access is enabled.
  constexpr int W = 11;
  const uint64_t my_pid = (uint64_t)getpid();
  std::vector<uint64_t> mine((size_t)k * W), all((size_t)N * k * W);
  for (int j = 0; j < k; ++j) {
    cudaIpcMemHandle_t h;
    BGX_CUDA(cudaIpcGetMemHandle(&h, local[j]));
    memcpy(&mine[(size_t)j * W], &h, 64);
    mine[(size_t)j * W + 8] = my_pid;
    mine[(size_t)j * W + 9] = (uint64_t)(uintptr_t)local[j];
    mine[(size_t)j * W + 10] = (uint64_t)c->device;
  }
  dist_allgather_host_u64(c, mine.data(), (size_t)k * W, all.data());
  for (int r = 0; r < N; ++r) {
    if (r == R) continue;
    for (int j = 0; j < k; ++j) {
      const uint64_t* hp = &all[((size_t)r * k + j) * W];
      if (hp[8] == my_pid) {
        const int dev = (int)hp[10];
        if (dev != c->device) {
          cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
          else BGX_CUDA(e);
        }
        peer[(size_t)j * N + r] = reinterpret_cast<void*>((uintptr_t)hp[9]);
        continue;
      }
      std::string key(reinterpret_cast<const char*>(hp), 64);
      key.push_back((char)r);
      auto it = c->dist.ipc_cache.find(key);
      if (it == c->dist.ipc_cache.end()) {
        cudaIpcMemHandle_t hr;
        memcpy(&hr, hp, 64);
        void* p = nullptr;
        BGX_CUDA(cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess));
        it = c->dist.ipc_cache.emplace(key, p).first;
      }
      peer[(size_t)j * N + r] = it->second;
    }
  }
}

void dist_map_peers(Context* c, void* local, void** peer) { dist_map_peers_n(c, &local, 1, peer); }

void dist_barrier(Context* c) {
  if (c->dist.nranks == 1) return;
  uint64_t one = 1;
  dist_allreduce_sum_host_u64(c, &one, 1);
}

Exchange dist_exchange_mode() {
  static const Exchange mode = [] {
    const char* e = getenv("BGX_EXCHANGE");
    if (e && std::string(e) == "nccl") return Exchange::NCCL;
    if (e && std::string(e) == "store") return Exchange::STORE;
    return Exchange::COPY;
  }();
  return mode;
}

namespace {
// All copies of a batch in ONE launch: block b works on copy b % n_copies, the blocks of a copy stride
// over it with 128-bit loads and stores (the stores go to the peer's memory over NVLink).  What NCCL's
// own kernels do for a send, without the handshake: the destination is mapped and known to be free.
struct CopyDesc {
  void* dst;
  const void* src;
  unsigned long long n;   // units: 16 bytes if wide, else 8
  int wide;
};
constexpr int kMaxCopies = 2 * 64;
struct CopyBatch {
  CopyDesc d[kMaxCopies];
  int n;
};
template <typename T>
__device__ __forceinline__ void copy_strided(T* __restrict__ dst, const T* __restrict__ src, unsigned long long n,
                                             unsigned long long first, unsigned long long stride) {
  unsigned long long i = first;
  // four independent transfers in flight per thread
  for (; i + 3 * stride < n; i += 4 * stride) {
    const T a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) dst[i] = src[i];
}
__global__ void __launch_bounds__(512) peer_copy_kernel(const CopyBatch* __restrict__ batch) {
  const int n = batch->n;
  const int which = blockIdx.x % n;
  const CopyDesc cd = batch->d[which];
  const unsigned blocks = (gridDim.x - which + n - 1) / n;   // blocks working on this copy
  const unsigned long long stride = (unsigned long long)blocks * blockDim.x;
  const unsigned long long first = (unsigned long long)(blockIdx.x / n) * blockDim.x + threadIdx.x;
  if (cd.wide) copy_strided(static_cast<uint4*>(cd.dst), static_cast<const uint4*>(cd.src), cd.n, first, stride);
  else copy_strided(static_cast<unsigned long long*>(cd.dst), static_cast<const unsigned long long*>(cd.src), cd.n, first, stride);
}
}  // namespace

void dist_peer_copies(Context* c, const std::vector<PeerCopy>& copies) {
  static const bool use_ce = [] { const char* e = getenv("BGX_PEER_COPY"); return e && std::string(e) == "ce"; }();  // A/B hook
  if (!use_ce) {
    // SM copy kernel (default): measured against the copy engines, which move ~110 GB/s per engine and
    // stream here (3 peers: 325 GB/s) -- see profiles/r2h_*
    cudaStream_t s = c->stream;
    CopyBatch h;
    h.n = 0;
    std::vector<PeerCopy> odd;   // not even 8-byte shaped: plain async copies
    for (const PeerCopy& pc : copies) {
      if (!pc.bytes) continue;
      const uintptr_t shape = (uintptr_t)pc.dst | (uintptr_t)pc.src | pc.bytes;
      if ((shape & 7) || h.n == kMaxCopies) { odd.push_back(pc); continue; }
      h.d[h.n].dst = pc.dst;
      h.d[h.n].src = pc.src;
      h.d[h.n].wide = (shape & 15) == 0;
      h.d[h.n].n = pc.bytes / (h.d[h.n].wide ? 16 : 8);
      ++h.n;
    }
    if (h.n) {
      DevBuf<CopyBatch> d(1, s);
      BGX_CUDA(cudaMemcpyAsync(d.p, &h, sizeof(CopyBatch), cudaMemcpyHostToDevice, s));
      BGX_CUDA(cudaStreamSynchronize(s));   // h lives on this stack frame
      const unsigned grid = (unsigned)(kNumSMs * 4 / h.n * h.n + (kNumSMs * 4 % h.n ? h.n : 0));
      KLAUNCH(peer_copy_kernel)<<<std::max<unsigned>(grid, (unsigned)h.n), 512, 0, s>>>(d.p);
      BGX_CUDA(cudaGetLastError());
    }
    for (const PeerCopy& pc : odd) BGX_CUDA(cudaMemcpyAsync(pc.dst, pc.src, pc.bytes, cudaMemcpyDefault, s));
    return;
  }
  const int N = c->dist.nranks;
  Dist& d = c->dist;
  if (d.copy_streams.empty()) {
    d.copy_streams.resize(N);
    d.copy_done.resize(N);
    for (int r = 0; r < N; ++r) {
      BGX_CUDA(cudaStreamCreateWithFlags(&d.copy_streams[r], cudaStreamNonBlocking));
      BGX_CUDA(cudaEventCreateWithFlags(&d.copy_done[r], cudaEventDisableTiming));
    }
    BGX_CUDA(cudaEventCreateWithFlags(&d.copy_go, cudaEventDisableTiming));
  }
  BGX_CUDA(cudaEventRecord(d.copy_go, c->stream));
  std::vector<char> used(N, 0);
  for (const PeerCopy& pc : copies) {
    if (!pc.bytes) continue;
    if (!used[pc.peer]) {
      BGX_CUDA(cudaStreamWaitEvent(d.copy_streams[pc.peer], d.copy_go, 0));
      used[pc.peer] = 1;
    }
    BGX_CUDA(cudaMemcpyAsync(pc.dst, pc.src, pc.bytes, cudaMemcpyDefault, d.copy_streams[pc.peer]));
  }
  for (int r = 0; r < N; ++r) {
    if (!used[r]) continue;
    BGX_CUDA(cudaEventRecord(d.copy_done[r], d.copy_streams[r]));
    BGX_CUDA(cudaStreamWaitEvent(c->stream, d.copy_done[r], 0));
  }
}

void dist_destroy(Context* c) {
  // Destroying a sharded context is collective: every rank first unmaps the peers' buffers, then all
  // ranks meet, and only then is device memory given back (a block must not be freed while a peer still
  // has it mapped).
  for (auto& kv : c->dist.ipc_cache) cudaIpcCloseMemHandle(kv.second);
  c->dist.ipc_cache.clear();
  if (c->dist.comm && c->dist.nranks > 1) {
    try { dist_barrier(c); } catch (...) {}
  }
  for (cudaStream_t st : c->dist.copy_streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (cudaEvent_t ev : c->dist.copy_done) cudaEventDestroy(ev);
  if (c->dist.copy_go) cudaEventDestroy(c->dist.copy_go);
  c->dist.copy_streams.clear();
  c->dist.copy_done.clear();
  c->dist.copy_go = nullptr;
  for (auto& kv : c->dist.ipc_cache) cudaIpcCloseMemHandle(kv.second);
  c->dist.ipc_cache.clear();
  if (c->dist.comm) {
    nccl().CommDestroy(comm_of(c));
    c->dist.comm = nullptr;
  }
  c->dist.nranks = 1;
  c->dist.rank = 0;
}

void dist_allreduce_sum_u64(Context* c, unsigned long long* buf, size_t n) {
  if (c->dist.nranks == 1 || n == 0) return;
  BGX_NCCL(nccl().AllReduce(buf, buf, n, ncclUint64, ncclSum, comm_of(c), c->stream));
}

void dist_allgather_bytes(Context* c, const void* send, void* recv, size_t bytes) {
  if (c->dist.nranks == 1) {
    if (send != recv && bytes) BGX_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return;
  }
  if (bytes == 0) return;
  BGX_NCCL(nccl().AllGather(send, recv, bytes, ncclUint8, comm_of(c), c->stream));
}

void dist_p2p_batch(Context* c, const std::vector<P2P>& sends, const std::vector<P2P>& recvs) {
  if (c->dist.nranks == 1) {
    // degenerate world: a send to self pairs with the recv of the same ordinal
    BGX_CHECK(sends.size() == recvs.size(), "dist_p2p_batch: unmatched self exchange");
    for (size_t i = 0; i < sends.size(); ++i) {
      BGX_CHECK(sends[i].bytes == recvs[i].bytes, "dist_p2p_batch: self exchange size mismatch");
      if (sends[i].bytes)
        BGX_CUDA(cudaMemcpyAsync(recvs[i].recv, sends[i].send, sends[i].bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    return;
  }
  // messages to self are plain device copies (pairing the i-th self send with the i-th self recv);
  // only peer traffic goes through NCCL
  {
    size_t ri = 0;
    for (const P2P& sd : sends) {
      if (sd.peer != c->dist.rank) continue;
      while (ri < recvs.size() && recvs[ri].peer != c->dist.rank) ++ri;
      BGX_CHECK(ri < recvs.size() && recvs[ri].bytes == sd.bytes, "dist_p2p_batch: unmatched self message");
      if (sd.bytes) BGX_CUDA(cudaMemcpyAsync(recvs[ri].recv, sd.send, sd.bytes, cudaMemcpyDeviceToDevice, c->stream));
      ++ri;
    }
  }
  BGX_NCCL(nccl().GroupStart());
  for (const P2P& sd : sends)
    if (sd.bytes && sd.peer != c->dist.rank) BGX_NCCL(nccl().Send(sd.send, sd.bytes, ncclUint8, sd.peer, comm_of(c), c->stream));
  for (const P2P& r : recvs)
    if (r.bytes && r.peer != c->dist.rank) BGX_NCCL(nccl().Recv(r.recv, r.bytes, ncclUint8, r.peer, comm_of(c), c->stream));
  BGX_NCCL(nccl().GroupEnd());
}

void dist_alltoallv(Context* c, const void* send, const uint64_t* send_off, const uint64_t* send_cnt, void* recv,
                    const uint64_t* recv_off, const uint64_t* recv_cnt, size_t elem_bytes) {
  std::vector<P2P> sends, recvs;
  for (int r = 0; r < c->dist.nranks; ++r) {
    P2P s, q;
    s.send = static_cast<const char*>(send) + send_off[r] * elem_bytes;
    s.bytes = send_cnt[r] * elem_bytes;
    s.peer = r;
    q.recv = static_cast<char*>(recv) + recv_off[r] * elem_bytes;
    q.bytes = recv_cnt[r] * elem_bytes;
    q.peer = r;
    sends.push_back(s);
    recvs.push_back(q);
  }
  dist_p2p_batch(c, sends, recvs);
}

void dist_allgather_host_u64(Context* c, const uint64_t* in, size_t n, uint64_t* out) {
  const int N = c->dist.nranks;
  if (N == 1) {
    memcpy(out, in, n * 8);
    return;
  }
  cudaStream_t s = c->stream;
  DevBuf<uint64_t> d_in(n, s), d_out(n * N, s);
  BGX_CUDA(cudaMemcpyAsync(d_in.p, in, n * 8, cudaMemcpyHostToDevice, s));
  dist_allgather_bytes(c, d_in.p, d_out.p, n * 8);
  BGX_CUDA(cudaMemcpyAsync(out, d_out.p, n * N * 8, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
}

void dist_allreduce_sum_host_u64(Context* c, uint64_t* v, size_t n) {
  if (c->dist.nranks == 1 || n == 0) return;
  cudaStream_t s = c->stream;
  DevBuf<unsigned long long> d(n, s);
  BGX_CUDA(cudaMemcpyAsync(d.p, v, n * 8, cudaMemcpyHostToDevice, s));
  dist_allreduce_sum_u64(c, d.p, n);
  BGX_CUDA(cudaMemcpyAsync(v, d.p, n * 8, cudaMemcpyDeviceToHost, s));
  BGX_CUDA(cudaStreamSynchronize(s));
}

}  // namespace bgx
